"""CPU execution of the CUDA thread bodies (csrc/ntt.cuh, csrc/dct.cuh via tests/emu/emu.cpp)
against the oracle.  Checks the kernel index math / lazy ranges / table conventions with no GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import ntt_perm

u64p = C.POINTER(C.c_uint64)
P = lambda a: a.ctypes.data_as(u64p)
BIG61 = [0x1fffffffffe00001, 0x1fffffffffc80001]  # BEHZ m_sk, gamma: 61-bit primes


def make_ring(E, n, qs, t=1 << 14):
    err = C.create_string_buffer(256)
    r = E.emu_ring_create(n, len(qs), (C.c_uint64 * len(qs))(*qs), t, err, 256)
    assert r, err.value
    return r


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("chain,rc", [("default", 0), ("default", 1), ("big61", 1)])
def test_emulated_ntt_matches_oracle(emu_lib, oracle_mod, n, chain, rc):
    qs = oracle_mod.default_coeff_modulus(4096) if chain == "default" else BIG61
    o = oracle_mod.Oracle(n, qs)
    r = make_ring(emu_lib, n, qs)
    rng = np.random.default_rng(n + rc)
    pm = ntt_perm(n)
    for j in range(len(qs)):
        q = np.uint64(qs[j])
        assert emu_lib.emu_psi(r, j) == o.psi(j) and emu_lib.emu_delta(r, j) == o.delta_mod(j)
        a = rng.integers(0, 2 ** 63, n, dtype=np.uint64) % q
        a[:4] = [0, 1, qs[j] - 1, qs[j] - 2]  # edge residues
        ref = a.copy()
        o.ntt_forward(j, ref)
        out = np.empty(n, dtype=np.uint64)
        assert emu_lib.emu_ntt_forward(r, j, P(a), P(out), 1, rc) == 0
        assert (out[pm] == ref).all()
        lazy = np.empty(n, dtype=np.uint64)
        emu_lib.emu_ntt_forward(r, j, P(a), P(lazy), 0, rc)
        assert ((lazy % q) == out).all()
        bound = 64 if rc == 0 else 8
        assert int(lazy.max()) < bound * qs[j]
        # inverse accepts values in [0, 2q)
        lz = out + rng.integers(0, 2, n).astype(np.uint64) * q
        back = np.empty(n, dtype=np.uint64)
        assert emu_lib.emu_ntt_inverse(r, j, P(lz), P(back), rc) == 0
        assert (back == a).all()
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
    ]
    for n, qs in bad:
        assert not emu_lib.emu_ring_create(n, len(qs), (C.c_uint64 * len(qs))(*qs), 1 << 14, err, 256)
        assert err.value
