"""CPU execution of the BEHZ kernel bodies (csrc/behz.cuh via tests/emu/emu.cpp) against
the oracle: base extension into the library's 55-bit auxiliary base, the transforms over B_sk, tensor, scaling
(fold_aux / reduce_u128 routes included) -- and the whole multiply sequenced as the library does it.  Checks the merged constants (m~^-1, t, q^-1, B^-1 folded into the
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
    L.emu_behz_bsk.restype = C.c_uint64
    L.emu_behz_bsk.argtypes = [C.c_void_p, C.c_int]
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


def aux_base(E, h):
    return [int(E.emu_behz_bsk(h, j)) for j in range(E.emu_behz_nb(h))]


@pytest.mark.parametrize("n,chain", [(1024, 2048), (1024, 4096), (2048, 8192)])
def test_aux_base_ntt_roundtrip_and_convolution(E, oracle_mod, n, chain):
    """the auxiliary primes (2^55 - j 2^15 + 1, csrc/behz_host.hpp) run the transforms of the q limbs"""
    o, r, h = make(E, oracle_mod, n, chain)
    bsk = aux_base(E, h)
    assert len(bsk) == len(o.bsk())                    # same limb count as SEAL's base
    assert len(set(bsk)) == len(bsk) and not set(bsk) & set(o.q)
    for p in bsk:
        assert p < 2 ** 55 and 2 ** 55 - p < 2 ** 26 and (p - 1) % (1 << 15) == 0 and (p - 1) % (1 << 18) != 0
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
    import pymodel
    o, r, h = make(E, oracle_mod, n, chain)
    bsk = aux_base(E, h)
    k, nb = o.k, len(bsk)
    # the per-coefficient bodies against the big-int model on THIS library's auxiliary base
    m = pymodel.Model(n, o.q, o.t)
    bz = pymodel.Behz(m, bsk)
    x = o.fill_uniform(0, 1, 1, 5)
    x[:2] = [0, o.q[0] - 1]
    e2 = np.empty(nb * n, dtype=np.uint64)
    E.emu_behz_extend(h, P(x), P(e2))
    want = bz.extend(m.from_rns([int(v) for v in x], 1)[0])
    assert [int(v) for v in e2] == [v for row in want for v in row]
    rng = np.random.default_rng(1)
    xq = np.concatenate([rng.integers(0, 2 ** 63, n, dtype=np.uint64) % np.uint64(q) for q in o.q])
    xb = np.concatenate([rng.integers(0, 2 ** 63, n, dtype=np.uint64) % np.uint64(p) for p in bsk])
    got = np.empty(k * n, dtype=np.uint64)
    E.emu_behz_scale(h, P(xq), P(xb), P(got))
    want = bz.scale_round([[int(v) for v in xq[i * n:(i + 1) * n]] for i in range(k)],
                          [[int(v) for v in xb[j * n:(j + 1) * n]] for j in range(nb)])
    assert [int(v) for v in got] == m.to_rns([want])
    # the whole multiply against the ORACLE, which uses SEAL's 61-bit base: the result does not depend on the base
    for s1, s2 in ((2, 2), (3, 2), (1, 4)):
        a, b = o.fill_uniform(3, 1, s1, 7), o.fill_uniform(9, 1, s2, 7)
        want = o.multiply(a, s1, b, s2)
        out = np.empty_like(want)
        assert E.emu_multiply(r, h, P(a), s1, P(b), s2, P(out)) == 0
        assert (want == out).all(), (n, chain, s1, s2)
    E.emu_behz_destroy(h)
    E.emu_ring_destroy(r)


def test_fold_aux_is_exact(E):
    """fold_aux (csrc/behz.cuh): z mod p for the auxiliary primes p = 2^55 - delta, z < 2^119"""
    E.emu_fold_aux.restype = C.c_uint64
    E.emu_fold_aux.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    primes = [(1 << 55) - (j << 15) + 1 for j in (1, 21, 333, 2047)]      # delta from 2^15 to just under 2^26
    rng = np.random.default_rng(55)
    M = (1 << 64) - 1
    for p in primes:
        zs = [0, 1, p - 1, p, p + 1, 2 * p, (1 << 119) - 1, (1 << 118), M, ((1 << 55) - 1) << 64 | M, (p << 55) | (p - 1)]
        zs += [(int(a) << 64 | int(b)) >> 9 for a, b in rng.integers(0, 1 << 63, (400, 2), dtype=np.uint64) * 2 + 1]
        zs += [(k * p - 1) for k in (1 << 60, (1 << 63) + 12345, (1 << 64) - 1)]
        for z in zs:
            assert z < 1 << 119
            assert E.emu_fold_aux(z & M, z >> 64, p) == z % p, (hex(p), hex(z))
