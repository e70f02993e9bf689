"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes mirror), against the CPU
oracle on the same seeded inputs -- bit-exact (integer residues), plus size-independent
properties at BASELINE.json's config-1 size.  Run on a B200:  pytest -m gpu"""
import json
import os

import numpy as np
import pytest

from conftest import ntt_perm

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gpu(pkg):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    pkg.build()
    return pkg


def make(gpu, oracle_mod, n, q=None):
    q = q or oracle_mod.default_coeff_modulus(n)
    return gpu.Context(n, q), oracle_mod.Oracle(n, q)


def dev_roundtrip(ctx, host):
    d = ctx.to_device(host)
    return d


def fetch(ctx, d, words):
    out = np.empty(words, dtype=np.uint64)
    ctx.download(out, d)
    return out


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384])
def test_ntt_forward_inverse_vs_oracle(gpu, oracle_mod, n):
    q = oracle_mod.default_coeff_modulus(n) if n >= 2048 else oracle_mod.default_coeff_modulus(4096)
    ctx, o = make(gpu, oracle_mod, n, q)
    k = ctx.k
    npoly = 3 * k  # three "polys" of k limbs, limb-cycled
    rng = np.random.default_rng(n)
    a = np.empty(npoly * n, dtype=np.uint64)
    for p in range(npoly):
        qq = np.uint64(q[p % k])
        a[p * n:(p + 1) * n] = rng.integers(0, 2 ** 63, n, dtype=np.uint64) % qq
        a[p * n:p * n + 3] = [0, q[p % k] - 1, 1]
    d = ctx.to_device(a)
    ctx.ntt_forward(d, npoly)
    got = fetch(ctx, d, a.size)
    pm = ntt_perm(n)
    for p in range(npoly):
        ref = a[p * n:(p + 1) * n].copy()
        o.ntt_forward(p % k, ref)
        assert (got[p * n:(p + 1) * n][pm] == ref).all(), (n, p)
    ctx.ntt_inverse(d, npoly)
    assert (fetch(ctx, d, a.size) == a).all()
    ctx.free(d)
    ctx.close()


def test_fill_uniform_and_checksum_match_oracle(gpu, oracle_mod):
    ctx, o = make(gpu, oracle_mod, 4096)
    n_ct = 9
    d = ctx.alloc(n_ct * ctx.ct_bytes())
    ctx.fill_uniform(d, 1234567, n_ct, 2, 0xABCDEF)
    got = fetch(ctx, d, n_ct * ctx.ct_words())
    ref = o.fill_uniform(1234567, n_ct, 2, 0xABCDEF)
    assert (got == ref).all()
    for j in range(ctx.k):
        assert got.reshape(n_ct, 2, ctx.k, ctx.n)[:, :, j].max() < ctx.q[j]
    assert ctx.checksum(d, got.size) == o.checksum(ref)
    assert ctx.checksum(d, 0) == 0
    ctx.free(d)
    ctx.close()


@pytest.mark.parametrize("n", [2048, 4096, 8192])
def test_evaluator_ops_vs_oracle(gpu, oracle_mod, n):
    ctx, o = make(gpu, oracle_mod, n)
    n_ct, W = 6, ctx.ct_words()
    a = o.fill_uniform(0, n_ct)
    b = o.fill_uniform(100, n_ct)
    b[:ctx.n] = a[:ctx.n]          # a - b == 0 on one limb
    a[ctx.n:ctx.n + 8] = 0         # zeros (negate keeps 0)
    enc = gpu.FractionalEncoder(n)
    # dense plaintexts (NTT route) and plaintexts with at most 8 coefficients, all +-1 (the signed-
    # rotation route of fhe_multiply_plain): 3 = 1 + x, 255 = 8 terms, 511 = 9 terms (dense route again)
    vals = [0.541196100, -1.847759065, 0.125, 3.0, 1 / 99, -128.0, 128.0, 2.0 ** -40,
            5.0, -3.0, 7.0, 0.5, 255.0, 511.0, -2.75, 1.0]
    plains = np.stack([enc.encode(v) for v in vals] + [np.zeros(n, dtype=np.uint64)])
    tab = gpu.PlainTable(ctx, plains)
    da, db = ctx.to_device(a), ctx.to_device(b)

    def check(op_gpu, op_ref):
        ctx.upload(da, a)
        op_gpu()
        got = fetch(ctx, da, a.size)
        ref = a.copy()
        for c in range(n_ct):
            op_ref(ref[c * W:(c + 1) * W], b[c * W:(c + 1) * W])
        assert (got == ref).all()

    check(lambda: ctx.add(da, db, n_ct), lambda x, y: o.add(x, y))
    check(lambda: ctx.sub(da, db, n_ct), lambda x, y: o.sub(x, y))
    check(lambda: ctx.negate(da, n_ct), lambda x, y: o.negate(x))
    for i, v in enumerate(vals):
        p = o.encode(v)
        assert (p == plains[i]).all()
        check(lambda: ctx.multiply_plain(da, n_ct, tab, i), lambda x, y: o.multiply_plain(x, p))
        check(lambda: ctx.add_plain(da, n_ct, tab, i), lambda x, y: o.add_plain(x, p))
        check(lambda: ctx.sub_plain(da, n_ct, tab, i), lambda x, y: o.sub_plain(x, p))
    with pytest.raises(gpu.FheError) as ei:   # SEAL: multiply_plain throws on a zero plaintext
        ctx.multiply_plain(da, n_ct, tab, len(vals))
    assert ei.value.code == -4
    with pytest.raises(gpu.FheError):
        ctx.multiply_plain(da, n_ct, tab, len(vals) + 1)
    ctx.add(da, db, 0)  # empty input is a no-op
    ctx.multiply_plain(da, 0, tab, 0)
    tab.close()
    ctx.free(da); ctx.free(db)
    ctx.close()


def oracle_dct(o, pkg, data, nblocks, quant):
    W = 64 * o.ct_words()
    for b in range(nblocks):
        blk = data[b * W:(b + 1) * W]
        o.encrypted_dct(blk)
        if quant is not None:
            o.quantize_fhe(blk, quant)
    return data


@pytest.mark.parametrize("n,nblocks,quant", [(4096, 2, "Y"), (4096, 1, None), (2048, 2, "UV"), (8192, 1, "Y"),
                                             (16384, 1, "Y")])
def test_dct_quantize_bit_exact_vs_oracle_graph(gpu, oracle_mod, n, nblocks, quant):
    ctx, o = make(gpu, oracle_mod, n)
    qt = {"Y": gpu.YQT, "UV": gpu.UVQT, None: None}[quant]
    plan = gpu.DctPlan(ctx, gpu.FractionalEncoder(n), qt)
    x = o.fill_uniform(5, 64 * nblocks)
    ref = oracle_dct(o, gpu, x.copy(), nblocks, qt)
    d_in, d_out = ctx.to_device(x), ctx.alloc(x.nbytes)
    gpu.encrypted_dct_quantize(ctx, plan, d_in, d_out, nblocks)     # out of place
    assert (fetch(ctx, d_out, x.size) == ref).all()
    assert (fetch(ctx, d_in, x.size) == x).all()                    # input untouched
    gpu.encrypted_dct_quantize(ctx, plan, d_in, d_in, nblocks)      # in place, as the reference
    assert (fetch(ctx, d_in, x.size) == ref).all()
    assert ctx.launch_count() > 0
    plan.close()
    ctx.free(d_in); ctx.free(d_out)
    ctx.close()


def test_dct_quantize_non_default_chain(gpu, oracle_mod):
    # any chain of primes 1 mod 2n works (SURVEY 0.1 caveat): three 55-bit primes at n=1024
    q = oracle_mod.default_coeff_modulus(8192)[:3]
    ctx, o = make(gpu, oracle_mod, 1024, q)
    plan = gpu.DctPlan(ctx, gpu.FractionalEncoder(1024), gpu.YQT)
    x = o.fill_uniform(7, 64 * 3, 2, 99)
    ref = oracle_dct(o, gpu, x.copy(), 3, gpu.YQT)
    d = ctx.to_device(x)
    gpu.encrypted_dct_quantize(ctx, plan, d, d, 3)
    assert (fetch(ctx, d, x.size) == ref).all()
    ctx.free(d)
    ctx.close()


def test_golden_checksums(gpu, oracle_mod):
    with open(os.path.join(HERE, "golden", "oracle_golden.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        ctx = gpu.Context(c["n"], c["q"], c["t"])
        plan = gpu.DctPlan(ctx, gpu.FractionalEncoder(c["n"], c["t"]), gpu.YQT if c["quant"] else None)
        words = 64 * c["nblocks"] * ctx.ct_words()
        d = ctx.alloc(words * 8)
        ctx.fill_uniform(d, c["first_ct"], 64 * c["nblocks"], 2, c["seed"])
        assert ctx.checksum(d, words) == int(c["input_checksum"], 16)
        gpu.encrypted_dct_quantize(ctx, plan, d, d, c["nblocks"])
        assert ctx.checksum(d, words) == int(c["output_checksum"], 16), c
        assert [int(v) for v in fetch(ctx, d, 4)] == c["first_words"]
        ctx.free(d)
        ctx.close()


def test_chunking_host_path_and_empty_input(gpu, oracle_mod):
    ctx, o = make(gpu, oracle_mod, 4096)
    plan = gpu.DctPlan(ctx, gpu.FractionalEncoder(4096), gpu.YQT)
    nb = 5
    x = o.fill_uniform(0, 64 * nb)
    d = ctx.to_device(x)
    gpu.encrypted_dct_quantize(ctx, plan, d, d, nb)
    whole = fetch(ctx, d, x.size)
    ctx.upload(d, x)
    ctx.set_dct_chunk_blocks(2)                 # ragged chunks 2+2+1
    gpu.encrypted_dct_quantize(ctx, plan, d, d, nb)
    assert (fetch(ctx, d, x.size) == whole).all()
    ctx.set_dct_chunk_blocks(0)
    gpu.encrypted_dct_quantize(ctx, plan, d, d, 0)   # zero blocks: no-op
    assert (fetch(ctx, d, x.size) == whole).all()
    # host-buffer entry point (pageable numpy buffers), then pinned
    out = np.zeros_like(x)
    gpu.encrypted_dct_quantize_host(ctx, plan, x, out, nb)
    assert (out == whole).all()
    hin, hout = gpu.host_alloc(x.nbytes), gpu.host_alloc(x.nbytes)
    import ctypes as C
    C.memmove(hin, x.ctypes.data, x.nbytes)
    gpu.encrypted_dct_quantize_host(ctx, plan, hin, hout, nb)
    got = np.frombuffer((C.c_uint64 * x.size).from_address(hout), dtype=np.uint64)
    assert (got == whole).all()
    # spot-check one block against the oracle
    W = ctx.block_words()
    ref = oracle_dct(o, gpu, x[3 * W:4 * W].copy(), 1, gpu.YQT)
    assert (whole[3 * W:4 * W] == ref).all()
    gpu.host_free(hin); gpu.host_free(hout)
    ctx.free(d)
    ctx.close()


def test_config1_full_size_properties(gpu, oracle_mod):
    """BASELINE config 1: 64x64 grayscale = 64 blocks (4096 ciphertexts), n=4096.
    Linearity DCT(a + b) == DCT(a) + DCT(b), homogeneity DCT(-a) == -DCT(a), sampled blocks
    bit-exact vs the oracle, on-device generation == oracle generation."""
    ctx, o = make(gpu, oracle_mod, 4096)
    plan = gpu.DctPlan(ctx, gpu.FractionalEncoder(4096), gpu.YQT)
    nb, n_ct = 64, 64 * 64
    words = n_ct * ctx.ct_words()
    da, db, dc = (ctx.alloc(words * 8) for _ in range(3))
    ctx.fill_uniform(da, 0, n_ct)
    ctx.fill_uniform(db, 1 << 20, n_ct)
    ctx.fill_uniform(dc, 0, n_ct)
    ctx.add(dc, db, n_ct)                                   # c = a + b
    for d in (da, db, dc):
        gpu.encrypted_dct_quantize(ctx, plan, d, d, nb)
    W = ctx.block_words()
    for blk in (0, 37, 63):                                 # sampled blocks vs the oracle
        x = o.fill_uniform(blk * 64, 64)
        ref = oracle_dct(o, gpu, x, 1, gpu.YQT)
        got = np.empty(W, dtype=np.uint64)
        ctx.download(got, da + blk * W * 8)
        assert (got == ref).all(), blk
    cs_c = ctx.checksum(dc, words)
    ctx.add(da, db, n_ct)                                   # DCT(a) + DCT(b)
    assert ctx.checksum(da, words) == cs_c
    # homogeneity: DCT(-b) + DCT(b) == 0
    ctx.fill_uniform(dc, 1 << 20, n_ct)
    ctx.negate(dc, n_ct)
    gpu.encrypted_dct_quantize(ctx, plan, dc, dc, nb)
    ctx.add(dc, db, n_ct)
    z = np.empty(words, dtype=np.uint64)
    ctx.download(z, dc)
    assert not z.any()
    for d in (da, db, dc):
        ctx.free(d)
    ctx.close()


@pytest.mark.parametrize("n", [2048, 8192])
def test_transparent_encryption_vs_oracle(gpu, oracle_mod, n):
    """fhe_encrypt_transparent == add_plain onto an all-zero ciphertext (the oracle's restatement of
    Evaluator::add_plain), and decrypts to the plaintext under any key."""
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q, 1 << 14, 64, 32)
    vals = [0.375, -17.25, 255.0, 1.0 / 3.0, -0.001]
    plains = np.stack([o.encode(v) for v in vals])
    d = ctx.alloc(len(vals) * ctx.ct_bytes(2))
    ctx.encrypt_transparent(plains, d)
    got = np.empty(len(vals) * ctx.ct_words(2), dtype=np.uint64)
    ctx.download(got, d)
    W = ctx.ct_words(2)
    sk, _ = o.keygen(4)
    for i, v in enumerate(vals):
        want = np.zeros(W, dtype=np.uint64)
        o.add_plain(want, plains[i])
        assert (got[i * W:(i + 1) * W] == want).all()
        assert abs(o.decode(o.decrypt(sk, got[i * W:(i + 1) * W].copy(), 2)) - v) < 1e-8
    # coefficients must be below t
    bad = plains[:1].copy()
    bad[0, 0] = 1 << 14
    with pytest.raises(gpu.FheError):
        ctx.encrypt_transparent(bad, d)
    ctx.free(d)
    ctx.close()


def test_device_block_cache(gpu, oracle_mod):
    """fhe_device_alloc / fhe_device_free keep freed blocks for exact-size reuse: same size -> same
    block back, a second free of a parked block and a foreign pointer are refused."""
    ctx = gpu.Context(2048, oracle_mod.default_coeff_modulus(2048))
    a = ctx.alloc(1 << 20)
    ctx.free(a)
    b = ctx.alloc(1 << 20)
    assert b == a                      # handed out again without touching the driver
    c = ctx.alloc(1 << 20)
    assert c != b
    ctx.free(b)
    with pytest.raises(gpu.FheError):
        ctx.free(b)                    # double free
    with pytest.raises(gpu.FheError):
        ctx.free(c + 256)              # not a block of this allocator
    ctx.free(c)
    ctx.close()


@pytest.mark.parametrize("n", [2048, 8192])
def test_sparse_multiply_plain_random_terms(gpu, oracle_mod, n):
    """The signed-rotation route of fhe_multiply_plain on random plaintexts with 1..8 coefficients in
    {+1, -1} at random exponents (including 0 and n-1: every wrap-around case), sizes 2 and 3, against
    the oracle's NTT-based multiply_plain; and 9 terms / a coefficient 2 take the dense route."""
    ctx, o = make(gpu, oracle_mod, n)
    t = int(ctx.t)
    rng = np.random.default_rng(n)
    plains = []
    for terms in (1, 2, 3, 5, 8, 8, 9):
        p = np.zeros(n, dtype=np.uint64)
        exps = rng.choice(n, size=terms, replace=False)
        exps[0] = [0, n - 1, 1, n // 2, n - 2, 7, 3][len(plains)]
        for e in exps:
            p[int(e)] = 1 if rng.integers(2) else t - 1
        plains.append(p)
    two = np.zeros(n, dtype=np.uint64)
    two[5] = 2
    plains.append(two)
    tab = gpu.PlainTable(ctx, np.stack(plains))
    for size in (2, 3):
        n_ct, W = 3, ctx.ct_words(size)
        a = o.fill_uniform(7, n_ct, size, 0x51DE)
        da = ctx.to_device(a)
        for i, p in enumerate(plains):
            ctx.upload(da, a)
            ctx.multiply_plain(da, n_ct, tab, i, size)
            got = fetch(ctx, da, a.size)
            ref = a.copy()
            for c in range(n_ct):
                o.multiply_plain(ref[c * W:(c + 1) * W], p, size)
            assert (got == ref).all(), (size, i)
        ctx.free(da)
    tab.close()
    ctx.close()
