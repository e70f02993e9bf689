"""CPU execution of the BEHZ kernel bodies (csrc/behz.cuh, ntt_wide.cuh via tests/emu/emu.cpp) against
the oracle: base extension, the 61-bit Harvey NTT over B_sk, tensor, scaling -- and the whole multiply
sequenced as the library does it.  Checks the merged constants (m~^-1, t, q^-1, B^-1 folded into the
conversion matrices) and the table conventions with no GPU."""
import ctypes as C

import numpy as np
import pytest

u64p = C.POINTER(C.c_uint64)
P = lambda a: a.ctypes.data_as(u64p)


@pytest.fixture(scope="module")
def E(emu_lib):
    L = emu_lib
    L.emu_behz_create.restype = C.c_void_p
    L.emu_behz_create.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.emu_behz_destroy.argtypes = [C.c_void_p]
    L.emu_behz_nb.restype = C.c_uint32
    L.emu_behz_nb.argtypes = [C.c_void_p]
    L.emu_nttw_forward.argtypes = [C.c_void_p, C.c_int, u64p, u64p]
    L.emu_nttw_inverse.argtypes = [C.c_void_p, C.c_int, u64p, u64p]
    L.emu_behz_extend.argtypes = [C.c_void_p, u64p, u64p]
    L.emu_behz_scale.argtypes = [C.c_void_p, u64p, u64p, u64p]
    L.emu_multiply.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_uint32, u64p, C.c_uint32, u64p]
    return L


def make(E, oracle_mod, n, chain_n):
    qs = oracle_mod.default_coeff_modulus(chain_n)
    o = oracle_mod.Oracle(n, qs, 1 << 14)
    err = C.create_string_buffer(256)
    r = E.emu_ring_create(n, len(qs), (C.c_uint64 * len(qs))(*qs), 1 << 14, err, 256)
    assert r, err.value
    h = E.emu_behz_create(r, err, 256)
    assert h, err.value
    return o, r, h


@pytest.mark.parametrize("n,chain", [(1024, 2048), (1024, 4096), (2048, 8192)])
def test_wide_ntt_roundtrip_and_convolution(E, oracle_mod, n, chain):
    o, r, h = make(E, oracle_mod, n, chain)
    bsk = o.bsk()
    assert E.emu_behz_nb(h) == len(bsk)
    rng = np.random.default_rng(n + chain)
    for j, p in enumerate(bsk):
        a = rng.integers(0, 2 ** 63, n, dtype=np.uint64) % np.uint64(p)
        b = rng.integers(0, 2 ** 63, n, dtype=np.uint64) % np.uint64(p)
        a[:3] = [0, p - 1, 1]
        x, y, z = (np.empty_like(a) for _ in range(3))
        E.emu_nttw_forward(h, j, P(a), P(x))
        assert x.max() < p
        E.emu_nttw_inverse(h, j, P(x), P(z))
        assert (z == a).all()
        # pointwise product in the NTT domain == negacyclic product (python ints, schoolbook on a sparse b)
        bs = np.zeros(n, dtype=np.uint64)
        bs[[0, 5, n - 1]] = [3, p - 2, 7]
        E.emu_nttw_forward(h, j, P(bs), P(y))
        prod = np.array([int(u) * int(v) % p for u, v in zip(x, y)], dtype=np.uint64)
        E.emu_nttw_inverse(h, j, P(prod), P(z))
        want = [0] * n
        for i, ai in enumerate(a):
            for d, bv in ((0, 3), (5, p - 2), (n - 1, 7)):
                t = i + d
                if t >= n:
                    want[t - n] = (want[t - n] - int(ai) * bv) % p
                else:
                    want[t] = (want[t] + int(ai) * bv) % p
        assert [int(v) for v in z] == want
    E.emu_behz_destroy(h)
    E.emu_ring_destroy(r)


@pytest.mark.parametrize("n,chain", [(1024, 2048), (1024, 4096), (2048, 8192), (1024, 16384)])
def test_extend_scale_and_multiply_bodies_match_oracle(E, oracle_mod, n, chain):
    o, r, h = make(E, oracle_mod, n, chain)
    k, nb = o.k, len(o.bsk())
    x = o.fill_uniform(0, 1, 1, 5)
    x[:2] = [0, o.q[0] - 1]
    e2 = np.empty(nb * n, dtype=np.uint64)
    E.emu_behz_extend(h, P(x), P(e2))
    assert (o.behz_extend(x) == e2).all()
    rng = np.random.default_rng(1)
    xq = np.concatenate([rng.integers(0, 2 ** 63, n, dtype=np.uint64) % np.uint64(q) for q in o.q])
    xb = np.concatenate([rng.integers(0, 2 ** 63, n, dtype=np.uint64) % np.uint64(p) for p in o.bsk()])
    got = np.empty(k * n, dtype=np.uint64)
    E.emu_behz_scale(h, P(xq), P(xb), P(got))
    assert (o.behz_scale_round(xq, xb) == got).all()
    for s1, s2 in ((2, 2), (3, 2), (1, 4)):
        a, b = o.fill_uniform(3, 1, s1, 7), o.fill_uniform(9, 1, s2, 7)
        want = o.multiply(a, s1, b, s2)
        out = np.empty_like(want)
        assert E.emu_multiply(r, h, P(a), s1, P(b), s2, P(out)) == 0
        assert (want == out).all(), (n, chain, s1, s2)
    E.emu_behz_destroy(h)
    E.emu_ring_destroy(r)


def test_fold61_is_exact_for_any_128_bit_value(E):
    """fold61 (csrc/behz.cuh): z mod p for the auxiliary primes p = 2^61 - delta, any z < 2^128"""
    E.emu_fold61.restype = C.c_uint64
    E.emu_fold61.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    primes = [0x1fffffffffe00001, 0x1fffffffffb40001, 0x1ffffffffe780001, 0x1ffffffffe4c0001]
    rng = np.random.default_rng(61)
    M = (1 << 64) - 1
    for p in primes:
        zs = [0, 1, p - 1, p, p + 1, 2 * p, (1 << 128) - 1, (1 << 127), M, M << 64, (p << 64) | (p - 1), ((1 << 61) - 1) << 64 | M]
        zs += [int(a) << 64 | int(b) for a, b in rng.integers(0, 1 << 63, (400, 2), dtype=np.uint64) * 2 + 1]
        zs += [(k * p - 1) % (1 << 128) for k in (1 << 64, (1 << 66) + 12345, (1 << 67) - 1)]
        for z in zs:
            assert E.emu_fold61(z & M, z >> 64, p) == z % p, (hex(p), hex(z))
