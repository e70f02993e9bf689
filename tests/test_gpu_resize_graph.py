"""GPU parity of the fused resize graphs (csrc/resize_api.inc: fhe_cubic, fhe_linear) against the
oracle's op-by-op restatement of Cubic / Linear (homo/fhe_resize.h:143-204): gathered taps, shared
fraction forms (t_div), both Cubic levels (sizes 2 -> 4 -> 6), at one limb and at the parameters of
BASELINE config 3 (n = 8192, k = 4, FractionalEncoder(100, 100)).  Bit-exact."""
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


@pytest.mark.parametrize("n,groups,div", [(2048, 2, 3), (8192, 2, 3), (4096, 3, 1)])
def test_fused_cubic_two_levels_and_linear_match_the_oracle(gpu, oracle_mod, n, groups, div):
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    enc = gpu.FractionalEncoder(n, ctx.t)
    plan = gpu.ResizePlan(ctx, enc)
    n_ct, W = groups * div, ctx.ct_words(2)
    # 4 * n_ct independent source ciphertexts, each in its OWN allocation (the taps are gathered by pointer)
    src = o.fill_uniform(0, 4 * n_ct, 2, 0xC0B1)
    d_src = [ctx.to_device(src[i * W:(i + 1) * W].copy()) for i in range(4 * n_ct)]
    t = o.fill_uniform(500, groups, 2, 0x7EA)
    d_t = ctx.to_device(t)
    ft = ctx.form(d_t, 2, groups)
    d_t2 = ctx.alloc(groups * ctx.ct_bytes(3))
    ctx.multiply_form(ft, 1, ft, 1, d_t2, groups)            # square(t); multiply(t, t) is the same ciphertext
    t2 = fetch(ctx, d_t2, groups * ctx.ct_words(3))
    for g in range(groups):
        assert (t2[g * ctx.ct_words(3):(g + 1) * ctx.ct_words(3)] == o.square(t[g * W:(g + 1) * W], 2)).all()
        assert (o.multiply(t[g * W:(g + 1) * W], 2, t[g * W:(g + 1) * W], 2) == o.square(t[g * W:(g + 1) * W], 2)).all()
    ft2 = ctx.form(d_t2, 3, groups)
    # level 1: size-2 taps -> size 4
    taps = [d_src[4 * i + j] for i in range(n_ct) for j in range(4)]
    W4, W6 = ctx.ct_words(4), ctx.ct_words(6)
    d_l1 = ctx.alloc(n_ct * ctx.ct_bytes(4))
    ctx.cubic(plan, taps, 2, ft, ft2, ft2, div, d_l1, n_ct)
    l1 = fetch(ctx, d_l1, n_ct * W4)
    want1 = []
    for i in range(n_ct):
        g = i // div
        c, rs = o.cubic(*[src[(4 * i + j) * W:(4 * i + j + 1) * W] for j in range(4)], [2, 2, 2, 2], t[g * W:(g + 1) * W])
        assert rs == 4
        want1.append(c)
        assert (l1[i * W4:(i + 1) * W4] == c).all(), ("level 1", n, i)
    # level 2: the size-4 results as taps (rotated, so the four taps differ) -> size 6
    taps2 = [d_l1 + ((i + j) % n_ct) * ctx.ct_bytes(4) for i in range(n_ct) for j in range(4)]
    d_l2 = ctx.alloc(n_ct * ctx.ct_bytes(6))
    ctx.cubic(plan, taps2, 4, ft, ft2, ft2, div, d_l2, n_ct)
    l2 = fetch(ctx, d_l2, n_ct * W6)
    for i in range(n_ct):
        g = i // div
        c, rs = o.cubic(*[want1[(i + j) % n_ct] for j in range(4)], [4, 4, 4, 4], t[g * W:(g + 1) * W])
        assert rs == 6
        assert (l2[i * W6:(i + 1) * W6] == c).all(), ("level 2", n, i)
    # the sources are untouched
    assert (fetch(ctx, d_src[1], W) == src[W:2 * W]).all()
    # Linear: (1 - t) A + t B, two levels (sizes 2 -> 3 -> 4)
    one = gpu.PlainTable(ctx, np.stack([enc.encode(1.0)]))
    d_1mt = ctx.to_device(t)
    ctx.negate(d_1mt, groups, 2)
    ctx.add_plain(d_1mt, groups, one, 0, 2)
    f1mt = ctx.form(d_1mt, 2, groups)
    W3 = ctx.ct_words(3)
    lt = [d_src[2 * i + j] for i in range(n_ct) for j in range(2)]
    d_lin = ctx.alloc(n_ct * ctx.ct_bytes(3))
    ctx.linear(lt, 2, f1mt, ft, div, d_lin, n_ct)
    lin = fetch(ctx, d_lin, n_ct * W3)
    wl = []
    for i in range(n_ct):
        g = i // div
        c, rs = o.linear(src[(2 * i) * W:(2 * i + 1) * W], 2, src[(2 * i + 1) * W:(2 * i + 2) * W], 2, t[g * W:(g + 1) * W])
        assert rs == 3
        wl.append(c)
        assert (lin[i * W3:(i + 1) * W3] == c).all(), ("linear", n, i)
    lt2 = [d_lin + ((i + j) % n_ct) * ctx.ct_bytes(3) for i in range(n_ct) for j in range(2)]
    d_lin2 = ctx.alloc(n_ct * ctx.ct_bytes(4))
    ctx.linear(lt2, 3, f1mt, ft, div, d_lin2, n_ct)
    lin2 = fetch(ctx, d_lin2, n_ct * W4)
    for i in range(n_ct):
        g = i // div
        c, rs = o.linear(wl[i], 3, wl[(i + 1) % n_ct], 3, t[g * W:(g + 1) * W])
        assert rs == 4
        assert (lin2[i * W4:(i + 1) * W4] == c).all(), ("linear 2", n, i)
    for f in (ft, ft2, f1mt):
        f.close()
    plan.close()
    one.close()
    ctx.close()


def test_resize_plan_refuses_plaintexts_that_are_not_short_sign_polynomials(gpu):
    import ctypes as C
    n = 2048
    ctx = gpu.Context(n)
    enc = gpu.FractionalEncoder(n, ctx.t)
    plains = np.ascontiguousarray(np.stack([enc.encode(v) for v in (3.0, 2.0, 5.0, 4.0, 0.3)]), dtype=np.uint64)   # 0.3: many terms
    h = C.c_void_p()
    rc = gpu.lib().fhe_resize_plan_create(ctx.h, plains.ctypes.data_as(gpu.u64p), C.byref(h))
    assert rc != 0 and b"short +-1 polynomials" in gpu.lib().fhe_last_error()
    ctx.close()
