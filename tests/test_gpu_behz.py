"""GPU parity tests of Evaluator::multiply / square / relinearize (csrc/behz.cuh, behz_api.inc)
through the C-ABI against the CPU oracle (oracle/fhe_oracle_behz.c), bit-exact, plus the committed
golden vectors produced by the big-int model where the ring size allows, and the semantic check
Dec(Eval(ct)) == Eval_plain(m).  Run on a B200:  pytest -m gpu"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(pkg):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    pkg.build()
    return pkg


def fetch(ctx, d, words):
    out = np.empty(words, dtype=np.uint64)
    ctx.download(out, d)
    return out


CASES = [  # n, chain source, (size_a, size_b), n_ct
    (4096, 4096, (2, 2), 3),
    (4096, 4096, (3, 2), 2),
    (4096, 4096, (4, 3), 1),
    (8192, 8192, (2, 2), 2),
    (8192, 8192, (4, 3), 1),
    (1024, 4096, (2, 2), 5),
    (2048, 2048, (2, 3), 2),
    (16384, 16384, (2, 2), 1),
    # the products the decode graphs form at BASELINE config 5's parameters (k = 4; fhe_decode.h:67-97,
    # 235, 239): x^8 * x, sin * cos, term * amplitude -- and the two largest of the bicubic graph
    (8192, 8192, (9, 2), 2),
    (8192, 8192, (11, 11), 1),
    (8192, 8192, (21, 2), 1),
    (8192, 8192, (4, 2), 3),
    (8192, 8192, (5, 5), 1),
]


@pytest.mark.parametrize("n,chain,sizes,n_ct", CASES)
def test_multiply_vs_oracle(gpu, oracle_mod, n, chain, sizes, n_ct):
    q = oracle_mod.default_coeff_modulus(chain)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    sa, sb = sizes
    a = o.fill_uniform(100, n_ct, sa, 0xBE42)
    b = o.fill_uniform(900, n_ct, sb, 0xBE43)
    # edge residues in the first ciphertext: 0 and q_j - 1 everywhere in poly 0
    av = a.reshape(n_ct, sa, ctx.k, n)
    av[0, 0, :, 0] = 0
    for j in range(ctx.k):
        av[0, 0, j, 1] = q[j] - 1
    da, db = ctx.to_device(a), ctx.to_device(b)
    dd = ctx.alloc(n_ct * ctx.ct_bytes(sa + sb - 1))
    ctx.multiply(da, sa, db, sb, dd, n_ct)
    got = fetch(ctx, dd, n_ct * ctx.ct_words(sa + sb - 1))
    wa, wb, wd = ctx.ct_words(sa), ctx.ct_words(sb), ctx.ct_words(sa + sb - 1)
    for i in range(n_ct):
        ref = o.multiply(a[i * wa:(i + 1) * wa], sa, b[i * wb:(i + 1) * wb], sb)
        assert (got[i * wd:(i + 1) * wd] == ref).all(), (n, sizes, i)
    # inputs untouched
    assert (fetch(ctx, da, a.size) == a).all() and (fetch(ctx, db, b.size) == b).all()
    for d in (da, db, dd):
        ctx.free(d)
    ctx.close()


@pytest.mark.parametrize("n,size,n_ct", [(4096, 2, 3), (4096, 3, 2), (8192, 2, 2), (4096, 5, 1), (8192, 3, 2), (8192, 5, 1)])
def test_square_vs_oracle_and_multiply(gpu, oracle_mod, n, size, n_ct):
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    a = o.fill_uniform(7, n_ct, size, 0x5151)
    da = ctx.to_device(a)
    wd = ctx.ct_words(2 * size - 1)
    d1, d2 = ctx.alloc(8 * n_ct * wd), ctx.alloc(8 * n_ct * wd)
    ctx.square(da, size, d1, n_ct)
    ctx.multiply(da, size, da, size, d2, n_ct)
    g1, g2 = fetch(ctx, d1, n_ct * wd), fetch(ctx, d2, n_ct * wd)
    assert (g1 == g2).all()
    wa = ctx.ct_words(size)
    for i in range(n_ct):
        assert (g1[i * wd:(i + 1) * wd] == o.square(a[i * wa:(i + 1) * wa], size)).all()
    for d in (da, d1, d2):
        ctx.free(d)
    ctx.close()


def test_multiply_semantics_on_real_encryptions(gpu, oracle_mod):
    n = 4096
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q, 1 << 14, 64, 32)
    sk, pk = o.keygen(41)
    vals = [(3.25, -1.5), (100.0, 0.125), (-7.0, -9.5)]
    a = np.concatenate([o.encrypt(pk, o.encode(x), 10 + i) for i, (x, _) in enumerate(vals)])
    b = np.concatenate([o.encrypt(pk, o.encode(y), 20 + i) for i, (_, y) in enumerate(vals)])
    da, db = ctx.to_device(a), ctx.to_device(b)
    dd = ctx.alloc(len(vals) * ctx.ct_bytes(3))
    ctx.multiply(da, 2, db, 2, dd, len(vals))
    got = fetch(ctx, dd, len(vals) * ctx.ct_words(3))
    W = ctx.ct_words(3)
    for i, (x, y) in enumerate(vals):
        plain, budget = o.decrypt(sk, got[i * W:(i + 1) * W].copy(), 3, True)
        assert abs(o.decode(plain) - x * y) < 1e-6 and budget > 20
    for d in (da, db, dd):
        ctx.free(d)
    ctx.close()


@pytest.mark.parametrize("n,dbc,size,n_ct", [(4096, 30, 3, 2), (4096, 16, 4, 2), (8192, 30, 3, 1), (2048, 20, 5, 1)])
def test_relinearize_vs_oracle(gpu, oracle_mod, n, dbc, size, n_ct):
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    sk, _ = o.keygen(77)
    count = size - 2
    evk = o.evk_gen(sk, dbc, count, 5)
    assert ctx.evk_digits(dbc) == o.evk_digits(dbc)
    keys = gpu.EvaluationKeys(ctx, evk, dbc, count)
    a = o.fill_uniform(55, n_ct, size, 0xE7)
    da = ctx.to_device(a)
    dd = ctx.alloc(n_ct * ctx.ct_bytes(2))
    ctx.relinearize(da, size, keys, dd, n_ct)
    got = fetch(ctx, dd, n_ct * ctx.ct_words(2))
    wa, wd = ctx.ct_words(size), ctx.ct_words(2)
    for i in range(n_ct):
        ref = o.relinearize(a[i * wa:(i + 1) * wa].copy(), size, evk, dbc, count)
        assert (got[i * wd:(i + 1) * wd] == ref).all(), (n, dbc, size, i)
    # not enough keys -> invalid argument, as seal::Evaluator::relinearize throws
    if size > 3:
        few = gpu.EvaluationKeys(ctx, evk[: evk.size // count], dbc, 1)
        with pytest.raises(gpu.FheError):
            ctx.relinearize(da, size, few, dd, n_ct)
        few.close()
    keys.close()
    ctx.free(da)
    ctx.free(dd)
    ctx.close()


def test_multiply_edge_cases(gpu, oracle_mod):
    """Empty batches, size-1 operands (a bare polynomial, as SEAL allows), the largest sizes of the
    decode graph (21 x 2 -> 22, fhe_decode.h:239), and rejected arguments."""
    n = 2048
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    a, b = o.fill_uniform(1, 1, 21, 3), o.fill_uniform(2, 1, 2, 4)
    da, db = ctx.to_device(a), ctx.to_device(b)
    dd = ctx.alloc(ctx.ct_bytes(22))
    ctx.multiply(da, 21, db, 2, dd, 0)            # n_ct = 0: nothing to do
    ctx.multiply(da, 21, db, 2, dd, 1)
    assert (fetch(ctx, dd, ctx.ct_words(22)) == o.multiply(a, 21, b, 2)).all()
    ctx.multiply(da, 1, db, 2, dd, 1)             # size 1 x size 2 -> size 2
    assert (fetch(ctx, dd, ctx.ct_words(2)) == o.multiply(a[: ctx.ct_words(1)].copy(), 1, b, 2)).all()
    for bad in (lambda: ctx.multiply(da, 0, db, 2, dd, 1), lambda: ctx.multiply(da, 33, db, 2, dd, 1),
                lambda: ctx.multiply(0, 2, db, 2, dd, 1), lambda: ctx.multiply(da, 2, db, 2, 0, 1),
                lambda: ctx.square(0, 2, dd, 1)):
        with pytest.raises(gpu.FheError):
            bad()
    for d in (da, db, dd):
        ctx.free(d)
    ctx.close()


def test_multiply_then_relinearize_decrypts(gpu, oracle_mod):
    n, dbc = 4096, 30
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q, 1 << 14, 64, 32)
    sk, pk = o.keygen(3)
    keys = gpu.EvaluationKeys(ctx, o.evk_gen(sk, dbc, 1, 9), dbc, 1)
    a, b = o.encrypt(pk, o.encode(12.5), 1), o.encrypt(pk, o.encode(-3.0), 2)
    da, db = ctx.to_device(a), ctx.to_device(b)
    d3, d2 = ctx.alloc(ctx.ct_bytes(3)), ctx.alloc(ctx.ct_bytes(2))
    ctx.multiply(da, 2, db, 2, d3)
    ctx.relinearize(d3, 3, keys, d2)
    got = fetch(ctx, d2, ctx.ct_words(2))
    plain, budget = o.decrypt(sk, got, 2, True)
    assert abs(o.decode(plain) + 37.5) < 1e-6 and budget > 20
    keys.close()
    for d in (da, db, d3, d2):
        ctx.free(d)
    ctx.close()


@pytest.mark.parametrize("n,sa,sb,groups,div", [(4096, 2, 3, 4, 3), (8192, 4, 3, 2, 3), (8192, 2, 2, 5, 1), (2048, 2, 2, 3, 2)])
def test_multiply_form_shares_operands_and_matches_multiply(gpu, oracle_mod, n, sa, sb, groups, div):
    """fhe_form_create / fhe_multiply_form: the transformed operand is built once and serves `div`
    consecutive products (the three colour channels share the fraction powers in Cubic); results are
    the oracle's multiply bit for bit, and a form times itself is square."""
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    n_ct = groups * div
    a = o.fill_uniform(10, n_ct, sa, 0xF0A1)
    b = o.fill_uniform(50, groups, sb, 0xF0B2)
    da, db = ctx.to_device(a), ctx.to_device(b)
    fa, fb = ctx.form(da, sa, n_ct), ctx.form(db, sb, groups)
    sd = sa + sb - 1
    dd = ctx.alloc(n_ct * ctx.ct_bytes(sd))
    ctx.multiply_form(fa, 1, fb, div, dd, n_ct)
    got = fetch(ctx, dd, n_ct * ctx.ct_words(sd))
    wa, wb, wd = ctx.ct_words(sa), ctx.ct_words(sb), ctx.ct_words(sd)
    for i in range(n_ct):
        g = i // div
        ref = o.multiply(a[i * wa:(i + 1) * wa], sa, b[g * wb:(g + 1) * wb], sb)
        assert (got[i * wd:(i + 1) * wd] == ref).all(), (n, i)
    # operand order does not matter; a form times itself is Evaluator::square
    ctx.multiply_form(fb, div, fa, 1, dd, n_ct)
    assert (fetch(ctx, dd, n_ct * ctx.ct_words(sd)) == got).all()
    d2 = ctx.alloc(groups * ctx.ct_bytes(2 * sb - 1))
    ctx.multiply_form(fb, 1, fb, 1, d2, groups)
    sq = fetch(ctx, d2, groups * ctx.ct_words(2 * sb - 1))
    w2 = ctx.ct_words(2 * sb - 1)
    for g in range(groups):
        assert (sq[g * w2:(g + 1) * w2] == o.square(b[g * wb:(g + 1) * wb], sb)).all()
    if div > 1:
        with pytest.raises(gpu.FheError):
            ctx.multiply_form(fa, 1, fb, 1, dd, n_ct)      # fb has too few ciphertexts for divisor 1
    fa.close(); fb.close()
    for d in (da, db, dd, d2):
        ctx.free(d)
    ctx.close()
