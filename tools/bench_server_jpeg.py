#!/usr/bin/env python
"""File-to-file throughput of server_jpeg_b200 (the streaming pipeline of server/stream_b200.hpp) on a synthetic
ciphertext stream in the wire format of host/seal_wire_b200.hpp: what a deployment behind the reference's
client/server split sees, record parser and file system included.

    python tools/bench_server_jpeg.py [--tiles 64] [--chunk 16] [--dir /dev/shm]

Prints one JSON line: tiles/s wall clock, the device time of the kernels (sum of the CSV values the server
prints), and the bytes read and written.  The input is uniform residues below the smallest prime (any
ciphertext is as good as another for timing)."""
import argparse
import json
import os
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib  # noqa: E402

pkg = importlib.import_module("fully-homomorphic-image-processing_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--tiles", type=int, default=64)
ap.add_argument("--chunk", type=int, default=16)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--dir", default="/dev/shm")
args = ap.parse_args()

n = args.n
q = pkg.coeff_modulus_128(n)
k = len(q)
exe = os.path.join(ROOT, pkg.__name__, "server", "server_jpeg_b200")
if not os.path.exists(exe):
    raise SystemExit("build first: python -c 'import __graft_entry__ as g; g.build()'")
work = os.path.join(args.dir, "fhe_b200_server_bench")
os.makedirs(os.path.join(work, "keys"), exist_ok=True)
os.makedirs(os.path.join(work, "image"), exist_ok=True)
side = int(round((args.tiles * 64) ** 0.5))
w = 8 * max(1, side // 8)
h = args.tiles * 64 // w
assert w * h == args.tiles * 64, "choose a tile count whose pixel count factors as 8a x b"
with open(os.path.join(work, "keys", "params.txt"), "w") as f:
    f.write(f"{w} {h} 3 {1 << 14}\n")
rng = np.random.default_rng(1)
rec = np.zeros((2 * k, n + 1), dtype="<u8")
head = struct.pack("<4Q", 1, 2, 3, 4) + struct.pack("<4i", 2, 2, n + 1, k)
t0 = time.time()
with open(os.path.join(work, "image", "nothingpersonnel.txt"), "wb") as f:
    base = rng.integers(0, min(q), size=(2 * k, n), dtype=np.uint64)
    for i in range(args.tiles * 192):
        rec[:, :n] = (base + np.uint64(i)) % np.uint64(min(q))
        f.write(head)
        f.write(rec.tobytes())
gen_s = time.time() - t0
in_bytes = os.path.getsize(os.path.join(work, "image", "nothingpersonnel.txt"))
t0 = time.time()
r = subprocess.run([exe, "--cmod", str(n), "--pmod", str(1 << 14), "--chunk", str(args.chunk), "-v"], cwd=work, capture_output=True, text=True)
wall = time.time() - t0
if r.returncode != 0:
    raise SystemExit("server failed: " + r.stderr[-500:])
dev_ms = 0.0
for line in r.stdout.splitlines():
    parts = [p for p in line.split(",") if p]
    if parts and parts[0] in ("RGBYCC", "DCT"):
        dev_ms += sum(float(x) for x in parts[1:])
out_bytes = os.path.getsize(os.path.join(work, "image", "zoop.txt"))
print(json.dumps({"what": "server_jpeg_b200 file to file (wire records parsed and written on helper threads, copies beside the kernels)",
                  "tiles": args.tiles, "chunk_tiles": args.chunk, "n": n, "k": k, "wall_s": wall, "tiles_per_s": args.tiles / wall,
                  "channel_blocks_per_s": 3 * args.tiles / wall, "kernel_ms_total": dev_ms, "kernel_share_of_wall": dev_ms * 1e-3 / wall,
                  "bytes_read": in_bytes, "bytes_written": out_bytes, "read_GBps": in_bytes / wall / 1e9,
                  "input_generation_s": gen_s, "dir": args.dir, "includes": "process start, context and plan creation"}))
for p in ("image/nothingpersonnel.txt", "image/zoop.txt"):
    os.remove(os.path.join(work, p))
