"""GPU parity of the fused server_jpeg tile step (csrc/jpeg_api.inc: rgb_to_ycc_fhe in the forward
NTT epilogue + encrypted_dct per channel [+ quantize_fhe]) against the oracle evaluating the
reference graph op by op: rgb_to_ycc_fhe per pixel (fhe_image.h:310-325), encrypted_dct per channel
(:196-288), quantize_fhe (:294-305).  Bit-exact.  Run on a B200:  pytest -m gpu"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(pkg):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    pkg.build()
    return pkg


def oracle_tiles(o, pkg, x, ntiles, quant):
    W = o.ct_words(2)
    t = x.reshape(ntiles, 3, 64, W).copy()
    for ti in range(ntiles):
        for px in range(64):
            r, g, b = (np.ascontiguousarray(t[ti, c, px]) for c in range(3))
            o.rgb_to_ycc_fhe(r, g, b)
            t[ti, 0, px], t[ti, 1, px], t[ti, 2, px] = r, g, b
        for c in range(3):
            blk = np.ascontiguousarray(t[ti, c]).reshape(-1)
            o.encrypted_dct(blk)
            if quant:
                o.quantize_fhe(blk, pkg.YQT if c == 0 else pkg.UVQT)
            t[ti, c] = blk.reshape(64, W)
    return t.reshape(-1)


@pytest.mark.parametrize("n,ntiles,quant,inplace", [(4096, 2, False, True), (4096, 1, True, False), (2048, 1, True, True),
                                                    (8192, 1, False, False)])
def test_fused_rgb_tile_vs_oracle(gpu, oracle_mod, n, ntiles, quant, inplace):
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    enc = gpu.FractionalEncoder(n, ctx.t)
    plan = gpu.JpegPlan(ctx, enc, gpu.YQT if quant else None, gpu.UVQT if quant else None)
    x = o.fill_uniform(17, ntiles * 192, 2, 0x7E9)
    d_in = ctx.to_device(x)
    d_out = d_in if inplace else ctx.alloc(x.nbytes)
    gpu.jpeg_rgb_tiles(ctx, plan, d_in, d_out, ntiles)
    got = np.empty_like(x)
    ctx.download(got, d_out)
    want = oracle_tiles(o, gpu, x, ntiles, quant)
    assert (got == want).all()
    if not inplace:
        chk = np.empty_like(x)
        ctx.download(chk, d_in)
        assert (chk == x).all()
        ctx.free(d_out)
    ctx.free(d_in)
    plan.close()
    ctx.close()


def test_fused_tile_equals_unfused_device_ops(gpu, oracle_mod):
    """The fused tile call == colour conversion through the Evaluator ops (multiply_plain / add / sub /
    sub_plain on arrays of pixels) followed by the fused DCT, all on the device -- at a size the
    oracle would need minutes for (8 tiles)."""
    n, ntiles = 4096, 8
    ctx = gpu.Context(n)
    enc = gpu.FractionalEncoder(n, ctx.t)
    plan = gpu.JpegPlan(ctx, enc)
    dplan = gpu.DctPlan(ctx, enc)
    tab = gpu.PlainTable(ctx, np.stack([enc.encode(c) for c in gpu.COLOR_CONSTANTS] + [enc.encode(128.0)]))
    W = ctx.ct_words(2)
    words = ntiles * 192 * W
    d_in = ctx.alloc(8 * words)
    ctx.fill_uniform(d_in, 5000, ntiles * 192, 2, 0xC0FFEE)
    d_fused = ctx.alloc(8 * words)
    gpu.jpeg_rgb_tiles(ctx, plan, d_in, d_fused, ntiles)
    cs_fused = ctx.checksum(d_fused, words)
    # unfused: per tile, channel arrays of 64 pixels
    import ctypes as C
    lib = gpu.lib()
    d_ref = ctx.alloc(8 * words)
    tmp = [ctx.alloc(8 * 64 * W) for _ in range(2)]

    def copy(dst, src, nwords):
        gpu._check(lib.fhe_device_copy(ctx.h, C.c_void_p(dst), C.c_void_p(src), 8 * nwords))

    for t in range(ntiles):
        base_in, base_out = d_in + 8 * t * 192 * W, d_ref + 8 * t * 192 * W
        R, G, B = (base_in + 8 * c * 64 * W for c in range(3))
        for ch, (cr, cg, cb, sg, sb) in enumerate([(0, 1, 2, +1, +1), (3, 4, 5, -1, +1), (6, 7, 8, -1, -1)]):
            acc = base_out + 8 * ch * 64 * W
            copy(acc, R, 64 * W)
            ctx.multiply_plain(acc, 64, tab, cr)
            for src, ci, sgn in ((G, cg, sg), (B, cb, sb)):
                copy(tmp[0], src, 64 * W)
                ctx.multiply_plain(tmp[0], 64, tab, ci)
                (ctx.add if sgn > 0 else ctx.sub)(acc, tmp[0], 64)
            if ch == 0:
                ctx.sub_plain(acc, 64, tab, 9)
            gpu.encrypted_dct_quantize(ctx, dplan, acc, acc, 1)
    assert ctx.checksum(d_ref, words) == cs_fused
    for d in [d_in, d_fused, d_ref] + tmp:
        ctx.free(d)
    tab.close(); plan.close(); dplan.close()
    ctx.close()
