#!/usr/bin/env python
"""A DCT+quantize call large enough that every CTA of the TMA-staged pointwise kernel walks several
units through its 3-stage ring (24 blocks -> 8 units per CTA), checked against the oracle.  Meant
to be run under compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_dct.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
pkg = importlib.import_module("fully-homomorphic-image-processing_b200")
import oracle_py  # noqa: E402  (the checker)

n, nblocks = 4096, 24
ctx = pkg.Context(n, None, pkg.PLAIN_MODULUS, device=0)
o = oracle_py.Oracle(n, ctx.q, ctx.t)
plan = pkg.DctPlan(ctx, pkg.FractionalEncoder(n, ctx.t), pkg.YQT)
x = o.fill_uniform(0, 64 * nblocks)
d = ctx.to_device(x)
pkg.encrypted_dct_quantize(ctx, plan, d, d, nblocks)
got = np.empty_like(x)
ctx.download(got, d)
W = ctx.block_words()
for b in range(nblocks):
    blk = x[b * W:(b + 1) * W]
    o.encrypted_dct(blk)
    o.quantize_fhe(blk, pkg.YQT)
assert (got == x).all()
print(f"sanitize_dct ok: {nblocks} blocks bit-exact vs oracle, {ctx.launch_count()} launches")
