"""Asynchronous host <-> device transfers of the C-ABI (include/fhe_b200.h: fhe_upload_async,
fhe_download_async, fhe_transfer_fence, fhe_upload_fence) -- the plumbing of the streaming servers
(server/stream_b200.hpp) and of bench.py's end-to-end legs.  More transfers than the library's event ring
holds, kernels between them, and the two host-side guarantees: after fhe_upload_fence(ctx, 0) the pinned
source of every upload may be overwritten; after fhe_transfer_fence(ctx, 0) every download has landed."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(pkg):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    pkg.build()
    return pkg


def pinned(pkg, words):
    p = C.c_void_p()
    pkg._check(pkg.lib().fhe_host_alloc(words * 8, C.byref(p)))
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(words,))


def test_async_transfers_round_trip_through_kernels(gpu, oracle_mod):
    n = 4096
    q = oracle_mod.default_coeff_modulus(n)
    ctx, o = gpu.Context(n, q), oracle_mod.Oracle(n, q)
    L, W, rounds = gpu.lib(), ctx.ct_words(2), 20           # 20 > the 8-deep event rings
    n_ct = 4
    src = [pinned(gpu, n_ct * W) for _ in range(rounds)]
    dst = [pinned(gpu, n_ct * W) for _ in range(rounds)]
    dev = [ctx.alloc(n_ct * W * 8) for _ in range(rounds)]
    want = []
    for i, (_, a) in enumerate(src):
        a[:] = o.fill_uniform(100 * i, n_ct, 2, 0xA5)
        neg = a.copy()
        for c in range(n_ct):
            o.negate(neg[c * W:(c + 1) * W], 2)
        want.append(neg)
    try:
        for i in range(rounds):
            gpu._check(L.fhe_upload_async(ctx.h, dev[i], src[i][0], n_ct * W * 8))
            gpu._check(L.fhe_transfer_fence(ctx.h, 8))      # the stream waits for the upload
            gpu._check(L.fhe_negate(ctx.h, dev[i], n_ct, 2))
            gpu._check(L.fhe_download_async(ctx.h, dst[i][0], dev[i], n_ct * W * 8))
        gpu._check(L.fhe_upload_fence(ctx.h, 0))
        for _, a in src:
            a[:] = 0                                         # the uploads are done: their sources are free
        gpu._check(L.fhe_transfer_fence(ctx.h, 0))
        for i in range(rounds):
            assert (dst[i][1] == want[i]).all(), i
        # keep_downloads: with two downloads in flight allowed, everything older has landed
        for i in range(rounds):
            dst[i][1][:] = 0
            gpu._check(L.fhe_download_async(ctx.h, dst[i][0], dev[i], n_ct * W * 8))
            gpu._check(L.fhe_transfer_fence(ctx.h, 2))
            if i >= 2:
                assert (dst[i - 2][1] == want[i - 2]).all(), i
        gpu._check(L.fhe_transfer_fence(ctx.h, 0))
        assert (dst[rounds - 1][1] == want[rounds - 1]).all()
    finally:
        ctx.sync()
        for p, _ in src + dst:
            L.fhe_host_free(p)
        for d in dev:
            ctx.free(d)
    assert L.fhe_upload_async(ctx.h, None, None, 8) != 0     # null buffers are refused
