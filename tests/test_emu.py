"""CPU execution of the CUDA thread bodies (csrc/ntt.cuh, csrc/dct.cuh via tests/emu/emu.cpp)
against the oracle.  Checks the kernel index math / lazy ranges / table conventions with no GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import ntt_perm

u64p = C.POINTER(C.c_uint64)
P = lambda a: a.ctypes.data_as(u64p)


def make_ring(E, n, qs, t=1 << 14):
    err = C.create_string_buffer(256)
    r = E.emu_ring_create(n, len(qs), (C.c_uint64 * len(qs))(*qs), t, err, 256)
    assert r, err.value
    return r


# a second chain next to SEAL's defaults: 50- and 41-bit primes, 1 mod 2^15
ALT = None


def alt_chain():
    global ALT
    if ALT is None:
        import sympy
        out, c = [], (1 << 50) // (1 << 15)
        while len(out) < 1:
            c -= 1
            if sympy.isprime(c * (1 << 15) + 1):
                out.append(c * (1 << 15) + 1)
        c = (1 << 41) // (1 << 15)
        while len(out) < 2:
            c -= 1
            if sympy.isprime(c * (1 << 15) + 1):
                out.append(c * (1 << 15) + 1)
        ALT = out
    return ALT


def test_lazy_arithmetic_ranges(emu_lib, oracle_mod):
    """mulc in [0, 5q), reduce63 in [0, 4q), canon exact -- for arguments up to 2^63 - 1."""
    for qs in (oracle_mod.default_coeff_modulus(4096), oracle_mod.default_coeff_modulus(2048), alt_chain()):
        r = make_ring(emu_lib, 4096, qs)
        rng = np.random.default_rng(7)
        for j, q in enumerate(qs):
            xs = [0, 1, q - 1, q, q + 1, 2 * q, 5 * q - 1, 2 ** 63 - 1, 2 ** 63 - q, (2 ** 63 // q) * q,
                  (2 ** 63 // q) * q - 1, 2 ** 32 - 1, 2 ** 32, 2 ** 62 + 2 ** 32 - 1]
            xs += [int(v) for v in rng.integers(0, 2 ** 63, 4000, dtype=np.uint64)]
            ws = [0, 1, q - 1, q // 2, q // 2 + 1] + [int(v) % q for v in rng.integers(0, 2 ** 63, 40, dtype=np.uint64)]
            for x in xs:
                y = emu_lib.emu_reduce63(r, j, x)
                assert y < 4 * q and y % q == x % q
                assert emu_lib.emu_canon(r, j, x) == x % q
            for x in xs[:200]:
                for w in ws:
                    y = emu_lib.emu_mulc(r, j, x, w)
                    assert y < 5 * q and y % q == (x * w) % q, (x, w, y)
        emu_lib.emu_ring_destroy(r)


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("chain", ["default", "alt"])
def test_emulated_ntt_matches_oracle(emu_lib, oracle_mod, n, chain):
    qs = oracle_mod.default_coeff_modulus(4096) if chain == "default" else alt_chain()
    o = oracle_mod.Oracle(n, qs)
    r = make_ring(emu_lib, n, qs)
    rng = np.random.default_rng(n)
    pm = ntt_perm(n)
    logn = n.bit_length() - 1
    for j in range(len(qs)):
        q = np.uint64(qs[j])
        assert emu_lib.emu_psi(r, j) == o.psi(j) and emu_lib.emu_delta(r, j) == o.delta_mod(j)
        a = rng.integers(0, 2 ** 63, n, dtype=np.uint64) % q
        a[:4] = [0, 1, qs[j] - 1, qs[j] - 2]  # edge residues
        ref = a.copy()
        o.ntt_forward(j, ref)
        out = np.empty(n, dtype=np.uint64)
        assert emu_lib.emu_ntt_forward(r, j, P(a), P(out), 2) == 0
        assert (out[pm] == ref).all()
        lazy = np.empty(n, dtype=np.uint64)
        emu_lib.emu_ntt_forward(r, j, P(a), P(lazy), 0)
        assert ((lazy % q) == out).all()
        assert int(lazy.max()) < (1 + 5 * logn) * qs[j]
        red = np.empty(n, dtype=np.uint64)
        emu_lib.emu_ntt_forward(r, j, P(a), P(red), 1)
        assert ((red % q) == out).all() and int(red.max()) < 4 * qs[j]
        # inverse accepts values in [0, 5q)
        lz = out + rng.integers(0, 5, n).astype(np.uint64) * q
        back = np.empty(n, dtype=np.uint64)
        assert emu_lib.emu_ntt_inverse(r, j, P(lz), P(back), 1) == 0
        assert (back == a).all()
        # unscaled variant: n * a mod q
        assert emu_lib.emu_ntt_inverse(r, j, P(lz), P(back), 0) == 0
        want = np.array([(int(v) * n) % qs[j] for v in a], dtype=np.uint64)
        assert (back == want).all()
    emu_lib.emu_ring_destroy(r)


def test_worst_case_inputs_stay_in_range(emu_lib, oracle_mod):
    """all-(q-1) and all-(5q-1) inputs drive every lazy bound to its maximum."""
    n, qs = 4096, oracle_mod.default_coeff_modulus(4096)
    o = oracle_mod.Oracle(n, qs)
    r = make_ring(emu_lib, n, qs)
    pm = ntt_perm(n)
    for j, q in enumerate(qs):
        a = np.full(n, q - 1, dtype=np.uint64)
        ref = a.copy()
        o.ntt_forward(j, ref)
        out = np.empty(n, dtype=np.uint64)
        emu_lib.emu_ntt_forward(r, j, P(a), P(out), 2)
        assert (out[pm] == ref).all()
        lz = np.full(n, 5 * q - 1, dtype=np.uint64)     # == q - 1 mod q at every NTT position
        ref2 = np.full(n, q - 1, dtype=np.uint64)[pm]   # any order: constant vector
        o.ntt_inverse(j, ref2)
        back = np.empty(n, dtype=np.uint64)
        emu_lib.emu_ntt_inverse(r, j, P(lz), P(back), 1)
        assert (back == ref2).all()
    emu_lib.emu_ring_destroy(r)


@pytest.mark.parametrize("n,quant", [(1024, True), (1024, False), (4096, True)])
def test_emulated_fused_dct_quantize_is_bit_exact(emu_lib, oracle_mod, pkg, n, quant):
    """The NTT-domain 64->64 map == the reference's op-by-op graph."""
    qs = oracle_mod.default_coeff_modulus(4096)
    o = oracle_mod.Oracle(n, qs)
    r = make_ring(emu_lib, n, qs)
    consts = np.stack([o.encode(c) for c in pkg.DCT_CONSTANTS])
    eighth = o.encode(0.125)
    qp = np.stack([o.encode(1 / v) for v in pkg.YQT]) if quant else None
    d = o.fill_uniform(0, 64)
    ref = d.copy()
    o.encrypted_dct(ref)
    if quant:
        o.quantize_fhe(ref, pkg.YQT)
    g = d.copy()
    assert emu_lib.emu_dct_quant(r, P(g), 1, P(consts), P(eighth), P(qp) if quant else None) == 0
    assert (g == ref).all()
    emu_lib.emu_ring_destroy(r)


def test_ring_rejects_bad_parameters(emu_lib):
    err = C.create_string_buffer(256)
    bad = [
        (1000, [0x7fffffff380001]),            # not a power of two
        (4096, [0x7fffffff380003]),            # not 1 mod 2n / not prime
        (4096, [0x7fffffff380001] * 2),        # duplicate prime
        (4096, [12289]),                       # <= t
        (4096, [0x1fffffffffe00001]),          # 61 bits: above SEAL 2.3's 60-bit limit for a coeff_modulus prime
    ]
    for n, qs in bad:
        assert not emu_lib.emu_ring_create(n, len(qs), (C.c_uint64 * len(qs))(*qs), 1 << 14, err, 256)
        assert err.value
