"""Any valid SEAL 2.3 coeff_modulus: a chain with a prime outside the window of the lazy arithmetic
([2^40, 2^55)) runs on the Harvey kernels of csrc/ntt_generic.cuh -- SEAL 2.3.0's 60-bit defaults, small
primes, mixed chains (the pinned SEAL fork's defaults are uncertain, so the chain is an
input).  Every Evaluator entry point and the fused resize graphs against the oracle, bit-exact; the
fused DCT says it needs the default-shaped primes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Q60 = [0xffffffffffc0001, 0xfffffffff840001]                 # 60-bit, 1 mod 2^18 (tests/golden/behz_golden.json)


def small_primes(n, count, bits=36):
    from sympy import isprime
    out, p = [], (1 << bits) + 1
    while len(out) < count:
        if (p - 1) % (2 * n) == 0 and isprime(p):
            out.append(p)
        p += 2 * n
    return out


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


def primes_60bit(count):
    """the largest primes below 2^60 that are 1 mod 2^18 (what SEAL 2.3.0 ships as its 60-bit defaults)"""
    from sympy import isprime
    out, p = [], (1 << 60) - (1 << 18) + 1
    while len(out) < count:
        if isprime(p):
            out.append(p)
        p -= 1 << 18
    return out


def chains():
    return [("60-bit", 4096, Q60), ("small", 2048, small_primes(2048, 2)), ("mixed", 4096, [0x7fffffff380001, Q60[0], small_primes(4096, 1)[0]]),
            ("60-bit k=4 n=8192", 8192, primes_60bit(4))]


@pytest.mark.parametrize("name,n,q", chains())
def test_evaluator_ops_on_a_generic_chain(gpu, oracle_mod, name, n, q):
    from sympy import isprime
    assert all(isprime(p) and (p - 1) % (2 * n) == 0 for p in q)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    enc = gpu.FractionalEncoder(n, ctx.t)
    W = ctx.ct_words(2)
    a, b = o.fill_uniform(0, 2, 2, 0x6E1), o.fill_uniform(10, 2, 2, 0x6E2)
    a.reshape(2, 2, ctx.k, n)[0, 0, :, 0] = 0
    for j in range(ctx.k):
        a.reshape(2, 2, ctx.k, n)[0, 0, j, 1] = q[j] - 1
    da, db = ctx.to_device(a), ctx.to_device(b)
    # transforms round trip, add / sub / negate, plaintext ops (dense and sparse multiply_plain)
    ctx.ntt_forward(da, 2 * 2 * ctx.k)
    ctx.ntt_inverse(da, 2 * 2 * ctx.k)
    assert (fetch(ctx, da, a.size) == a).all()
    tab = gpu.PlainTable(ctx, np.stack([enc.encode(0.587), enc.encode(3.0), enc.encode(-128.0)]))
    want = a.copy()
    for i in range(2):
        x = want[i * W:(i + 1) * W]
        o.add(x, b[i * W:(i + 1) * W])
        o.multiply_plain(x, o.encode(0.587))
        o.sub(x, b[i * W:(i + 1) * W])
        o.multiply_plain(x, o.encode(3.0))
        o.add_plain(x, o.encode(-128.0))
        o.negate(x)
    ctx.add(da, db, 2)
    ctx.multiply_plain(da, 2, tab, 0)
    ctx.sub(da, db, 2)
    ctx.multiply_plain(da, 2, tab, 1)
    ctx.add_plain(da, 2, tab, 2)
    ctx.negate(da, 2)
    assert (fetch(ctx, da, a.size) == want).all(), name
    # multiply, square, relinearize
    ctx.upload(da, a)
    d3, d2 = ctx.alloc(2 * ctx.ct_bytes(3)), ctx.alloc(2 * ctx.ct_bytes(2))
    ctx.multiply(da, 2, db, 2, d3, 2)
    prod = fetch(ctx, d3, 2 * ctx.ct_words(3))
    W3 = ctx.ct_words(3)
    for i in range(2):
        assert (prod[i * W3:(i + 1) * W3] == o.multiply(a[i * W:(i + 1) * W], 2, b[i * W:(i + 1) * W], 2)).all(), (name, "multiply", i)
    ctx.square(db, 2, d3, 2)
    sq = fetch(ctx, d3, 2 * W3)
    assert (sq[:W3] == o.square(b[:W], 2)).all(), (name, "square")
    sk, _ = o.keygen(4)
    dbc = 30
    evk = o.evk_gen(sk, dbc, 1, 13)
    keys = gpu.EvaluationKeys(ctx, evk, dbc, 1)
    ctx.relinearize(d3, 3, keys, d2, 2)
    rel = fetch(ctx, d2, 2 * W)
    assert (rel[:W] == o.relinearize(sq[:W3], 3, evk, dbc, 1)).all(), (name, "relinearize")
    # the fused Cubic on this chain
    try:
        plan = gpu.ResizePlan(ctx, enc)
    except gpu.FheError:
        plan = None
    if plan is not None:
        src = o.fill_uniform(50, 4, 2, 0x6E3)
        d_src = [ctx.to_device(src[i * W:(i + 1) * W].copy()) for i in range(4)]
        ft = ctx.form(db, 2, 1)
        ctx.multiply_form(ft, 1, ft, 1, d3, 1)
        ft2 = ctx.form(d3, 3, 1)
        d4 = ctx.alloc(ctx.ct_bytes(4))
        ctx.cubic(plan, d_src, 2, ft, ft2, ft2, 1, d4, 1)
        c, rs = o.cubic(*[src[i * W:(i + 1) * W] for i in range(4)], [2, 2, 2, 2], b[:W])
        assert rs == 4 and (fetch(ctx, d4, ctx.ct_words(4)) == c).all(), (name, "cubic")
        ft.close(); ft2.close(); plan.close()
    # the fused DCT needs the default-shaped primes and says so
    with pytest.raises(gpu.FheError, match="2\\^40, 2\\^55"):
        gpu.DctPlan(ctx, enc, gpu.YQT)
    keys.close(); tab.close(); ctx.close()
