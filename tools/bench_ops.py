#!/usr/bin/env python
"""bench_ops.py -- device-time of the Evaluator primitives the resize / decode configs are built
from (BASELINE.json configs 3 and 5: n=8192, k=4), batched over independent ciphertexts, through
the C-ABI.  Not the headline bench (bench.py); prints one JSON object per op for profiles/.

  python tools/bench_ops.py [--n 8192] [--batch 256]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    assert torch.cuda.is_available()
    pkg = importlib.import_module("fully-homomorphic-image-processing_b200")
    pkg.build()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = pkg.Context(args.n, None, pkg.PLAIN_MODULUS, device=0, stream=stream.cuda_stream)
    enc = pkg.FractionalEncoder(args.n, ctx.t)
    N, W = args.batch, ctx.ct_words(2)
    k = ctx.k

    def buf(size):
        return torch.empty(N * ctx.ct_words(size), dtype=torch.int64, device="cuda")

    a2, b2, a3, a4 = buf(2), buf(2), buf(3), buf(4)
    ctx.fill_uniform(a2.data_ptr(), 0, N, 2, 1)
    ctx.fill_uniform(b2.data_ptr(), 7 * N, N, 2, 2)
    ctx.fill_uniform(a3.data_ptr(), 0, N, 3, 3)
    ctx.fill_uniform(a4.data_ptr(), 0, N, 4, 4)
    d3, d4, d5, d6 = buf(3), buf(4), buf(5), buf(6)
    tab = pkg.PlainTable(ctx, np.stack([enc.encode(0.587)]))
    dbc = 30
    nd = ctx.evk_digits(dbc)
    rng = np.random.default_rng(1)
    evk = np.empty(2 * k * nd * 2 * k * args.n, dtype=np.uint64).reshape(-1, k, args.n)
    for j in range(k):
        evk[:, j, :] = rng.integers(0, ctx.q[j], size=(evk.shape[0], args.n), dtype=np.uint64)
    keys = pkg.EvaluationKeys(ctx, evk.reshape(-1), dbc, 2)

    ops = {
        "add (size 2)": (lambda: ctx.add(a2.data_ptr(), b2.data_ptr(), N, 2), 3 * 2),
        "multiply_plain (size 2)": (lambda: ctx.multiply_plain(a2.data_ptr(), N, tab, 0, 2), 2 * 2),
        "multiply 2x2 -> 3": (lambda: ctx.multiply(a2.data_ptr(), 2, b2.data_ptr(), 2, d3.data_ptr(), N), 2 + 2 + 3),
        "square 2 -> 3": (lambda: ctx.square(a2.data_ptr(), 2, d3.data_ptr(), N), 2 + 3),
        "multiply 3x2 -> 4": (lambda: ctx.multiply(a3.data_ptr(), 3, b2.data_ptr(), 2, d4.data_ptr(), N), 3 + 2 + 4),
        "multiply 4x3 -> 6": (lambda: ctx.multiply(a4.data_ptr(), 4, a3.data_ptr(), 3, d6.data_ptr(), N), 4 + 3 + 6),
        "relinearize 3 -> 2 (dbc 30)": (lambda: ctx.relinearize(a3.data_ptr(), 3, keys, b2.data_ptr(), N), 3 + 2),
        "relinearize 4 -> 2 (dbc 30)": (lambda: ctx.relinearize(a4.data_ptr(), 4, keys, b2.data_ptr(), N), 4 + 2),
    }
    poly_bytes = k * args.n * 8
    out = {"n": args.n, "k": k, "batch": N, "ops": {}}
    for name, (fn, polys_moved) in ops.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        l0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        alg = polys_moved * poly_bytes * N
        out["ops"][name] = {"ms_per_batch": ms, "ops_per_s": N / (ms * 1e-3),
                            "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
                            "launches_per_call": (ctx.launch_count() - l0) // args.reps}
    print(json.dumps(out))
    keys.close()
    tab.close()
    ctx.close()


if __name__ == "__main__":
    main()
