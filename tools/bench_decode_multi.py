#!/usr/bin/env python
"""BASELINE config 5 on N GPUs of one box: run-length pairs are independent, so each
GPU gets its own process running tools/bench_decode (one pair at a time, no collective) and the job
rate is the sum of the per-GPU rates; the wall time of the slowest process is reported beside it.
    python tools/bench_decode_multi.py --gpus 8 [--dbc 30] [--n 8192] [--positions 64] [--degree 12]"""
import argparse
import json
import os
import subprocess
import time

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--positions", type=int, default=64)
ap.add_argument("--degree", type=int, default=12)
ap.add_argument("--order", type=int, default=64)
ap.add_argument("--dbc", type=int, default=0, help="relinearise every product with this decomposition bit count (0 = as the reference)")
a = ap.parse_args()
exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_decode")
t0 = time.perf_counter()
procs = [subprocess.Popen([exe, str(a.n), str(a.positions), str(a.degree), str(a.order), str(a.dbc)],
                          env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(g)), stdout=subprocess.PIPE, text=True)
         for g in range(a.gpus)]
outs = [p.communicate()[0] for p in procs]
wall = time.perf_counter() - t0
assert all(p.returncode == 0 for p in procs), [p.returncode for p in procs]
per = [json.loads(o.strip().splitlines()[-1]) for o in outs]
secs = [r["seconds_per_run_length_pair"] for r in per]
print(json.dumps({"what": per[0]["what"], "n_gpus": a.gpus, "pairs_per_s": sum(1.0 / s for s in secs),
                  "seconds_per_pair_per_gpu": secs, "reference_graph_ct_products_per_s": sum(r["reference_graph_ct_products_per_s"] for r in per),
                  "out_ct_size": per[0]["out_ct_size"], "wall_s_including_setup": wall}))
