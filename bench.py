#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (encrypted 8x8 DCT + quantize blocks/s) and the two other
hot-path configurations (bicubic resize, homomorphic run-length decode) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 # headline: DCT+quantize (config 2)
  python bench.py --workload resize|decode ...                        # configs 3 and 5, own JSON line
  python bench.py --impl reference [--workload ...] ...               # CPU arm (oracle op-by-op graph)
  torchrun ... bench.py --gpus N [--ingest scatter] ...               # one rank per GPU, weak scaling

dct (default).  A "step" is one pass of encrypted_dct + quantize_fhe (fhe_image.h:196-305) over one
resident chunk of synthetic ciphertext blocks of BASELINE config 2 (server_jpeg DCT+quantize,
1024x1024 RGB = 49152 blocks, n=4096, k=2, t=2^14).  The 384 GiB image does not fit in HBM, so it
is streamed as chunks of --blocks blocks (default 8192 = 1/6 of the image, 64 GiB in + 64 GiB
out); inputs are generated on the device (seeded SplitMix64, uniform residues)
before the timed region.  The default line also carries the other two workloads at their full
configurations with fewer steps (keys "config3_resize", "config5_decode"; --no-side skips them).
resize.  A step is a band of --rows output rows of BASELINE config 3 (server_resize --bicubic,
512x512 -> 256x256 RGB, n=8192, k=4) through the row engine of host/fhe_resize_b200.hpp.
decode.  A step is one run-length pair of BASELINE config 5 (approximated_step, 64 positions, degree
12, order 64, n=8192, every product relinearised at dbc 30) through host/fhe_decode_b200.hpp.
Prints ONE JSON line (rank 0).
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RING, T_PLAIN = 4096, 1 << 14
IMAGE_BLOCKS = 3 * 16384  # 1024x1024 RGB
METRIC = "encrypted 8x8 DCT+quantize blocks/sec"
WORKLOAD = ("server_jpeg DCT+quantize, 1024x1024 RGB (3x16384 blocks), n=4096 k=2 t=2^14, "
            "streamed as HBM-resident chunks")
UNIT = "blocks/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sample(threads, blocks, min_seconds=0.0, keep_first=None):
    """Oracle (CPU restatement of the reference's op-by-op graph) on a bounded sample: `blocks`
    blocks per pass, repeated on fresh inputs until `min_seconds` of CPU work were timed.
    keep_first: a list that receives the oracle's OUTPUT of the first pass (blocks 0 .. blocks-1 of
    the workload), so the caller can check the GPU's output of the same blocks against it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    pkg = importlib.import_module("fully-homomorphic-image-processing_b200")
    o = oracle_py.Oracle(N_RING, None, T_PLAIN)
    secs, done = 0.0, 0
    while True:
        data = o.fill_uniform(64 * done, 64 * blocks)
        secs += o.bench_dct_quant(data, blocks, pkg.YQT, threads)
        if keep_first is not None and done == 0:
            keep_first.append(data)
        done += blocks
        if secs >= min_seconds:
            break
    return done / secs, secs, done


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path.  SEAL is absent from /root/reference (empty
    submodule), so this is the oracle port ("kind": "port"): encrypted_dct + quantize_fhe op by op
    with the reference's copies, blocks spread over all host threads.  Rank 0 only."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    blocks = 8 * threads                      # ~1 s of CPU work per step on the GPU box's host cores
    for _ in range(args.warmup):
        cpu_sample(threads, max(1, threads // 2))
    t = 0.0
    for _ in range(args.steps):
        _, s, _ = cpu_sample(threads, blocks)
        t += s
    value = blocks * args.steps / t
    sample = f"{blocks} blocks/step (64 size-2 cts each, n=4096 k=2), {threads} threads, op-by-op graph"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "blocks_per_step": blocks, "image_blocks": IMAGE_BLOCKS,
                   "note": "bounded CPU sample of the same workload on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=8192, help="blocks per step per GPU (resident chunk)")
    ap.add_argument("--e2e-blocks", type=int, default=256, help="blocks per end-to-end step (host buffers)")
    ap.add_argument("--chunk", type=int, default=0, help="internal chunk blocks (0 = one chunk)")
    ap.add_argument("--workload", default="dct", choices=["dct", "resize", "decode"],
                    help="dct: BASELINE config 2 (headline); resize: config 3; decode: config 5")
    ap.add_argument("--rows", type=int, default=8, help="resize: output rows per step per GPU")
    ap.add_argument("--no-side", action="store_true", help="dct: skip the resize / decode side measurements")
    ap.add_argument("--ingest", default="local", choices=["local", "scatter"],
                    help="dct: also measure the one-rank ingest path (NCCL scatter of inputs from rank 0, gather of results)")
    ap.add_argument("--ingest-blocks", type=int, default=512, help="scatter path: blocks per rank")
    ap.add_argument("--ingest-chunk", type=int, default=128, help="scatter path: blocks per NCCL message (1 GiB at n=4096)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-tile", action="store_true", help="skip the fused RGB-tile side measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.workload == "dct":
            run_reference(args, rank, world)
        elif rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_workloads
            print(json.dumps(bench_workloads.reference_resize(args) if args.workload == "resize" else bench_workloads.reference_decode(args)))
        return

    import torch
    import torch.distributed as dist
    args.warmup = max(args.warmup, 3)  # timing rules: at least 3 warm-up steps (reported as run)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("fully-homomorphic-image-processing_b200")
    pkg.build()

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    if args.workload != "dct":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_workloads
        run = bench_workloads.run_resize if args.workload == "resize" else bench_workloads.run_decode
        line = run(args, rank, world, local_rank, torch, dist, stream, ClockSampler, measured_peaks())
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return
    ctx = pkg.Context(N_RING, None, T_PLAIN, device=local_rank, stream=stream.cuda_stream)
    enc = pkg.FractionalEncoder(N_RING, T_PLAIN)
    plan = pkg.DctPlan(ctx, enc, pkg.YQT)
    if args.chunk:
        ctx.set_dct_chunk_blocks(args.chunk)

    # ---- resident inputs: this rank's shard of the (weak-scaled) job --------------------------
    B = args.blocks
    words = B * ctx.block_words()
    d_in = torch.empty(words, dtype=torch.int64, device="cuda")
    d_out = torch.empty(words, dtype=torch.int64, device="cuda")
    first_block = rank * B
    ctx.fill_uniform(d_in.data_ptr(), first_block * 64, B * 64)
    ctx.sync()

    def step():
        pkg.encrypted_dct_quantize(ctx, plan, d_in.data_ptr(), d_out.data_ptr(), B)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # per-kernel times (events between the three kernels), outside the timed region
    ctx.set_stage_timing(True)
    stage = [0.0, 0.0, 0.0]
    for _ in range(3):
        step()
        st = ctx.stage_times_ms()
        stage = [a + b / 3 for a, b in zip(stage, st)]
    ctx.set_stage_timing(False)

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * B * args.steps / (ms * 1e-3)

    # rank 0 keeps its first SPOT output blocks (blocks 0..SPOT-1 of the workload) for the bit-exact
    # check against the oracle in the cpu_baseline leg
    SPOT = min(4, B)
    spot = None
    if rank == 0 and not args.no_cpu:
        import numpy as np
        spot = np.empty(SPOT * ctx.block_words(), dtype=np.uint64)
        ctx.download(spot, d_out.data_ptr())

    # edge collective only: checksum of every rank's output
    cs = ctx.checksum(d_out.data_ptr(), words)
    cs_t = torch.tensor([cs & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
    if world > 1:
        # NCCL has no XOR reduction: gather the per-rank checksums and fold them on the host
        allcs = torch.empty(world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allcs, cs_t)
        folded = 0
        for v in allcs.tolist():
            folded ^= int(v)
        cs_t = torch.tensor([folded], dtype=torch.int64, device="cuda")

    # ---- side measurement (not the headline): the whole server_jpeg tile step, colour conversion
    # as an NTT-domain pass (fhe_jpeg_rgb_tiles), in place on the same resident buffer ------
    tile_info = None
    if not args.no_tile:
        jplan = pkg.JpegPlan(ctx, enc, pkg.YQT, pkg.UVQT)
        ntiles = B // 3
        for _ in range(2):
            pkg.jpeg_rgb_tiles(ctx, jplan, d_in.data_ptr(), d_in.data_ptr(), ntiles)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record(stream)
        for _ in range(3):
            pkg.jpeg_rgb_tiles(ctx, jplan, d_in.data_ptr(), d_in.data_ptr(), ntiles)
        t1e.record(stream)
        torch.cuda.synchronize()
        tms = t0e.elapsed_time(t1e) / 3
        tile_info = {"what": "rgb_to_ycc_fhe + encrypted_dct + quantize_fhe per 8x8 RGB tile (3 channel blocks), "
                             "fused, per GPU", "tiles_per_s": ntiles / (tms * 1e-3),
                     "channel_blocks_per_s": 3 * ntiles / (tms * 1e-3), "ms": tms, "tiles": ntiles}
        jplan.close()

    del d_in, d_out
    torch.cuda.empty_cache()
    # ---- one-rank ingest (north_star: "NCCL over NVLink only to scatter input ciphertexts and gather
    # results"): the whole input stream is resident on rank 0, chunks travel over NCCL point-to-point,
    # every rank computes its share, results return to rank 0's output stream (ingest.py)
    ingest_info = None
    if args.ingest == "scatter":
        ingest = importlib.import_module("fully-homomorphic-image-processing_b200.ingest")
        total, bw = world * args.ingest_blocks, ctx.block_words()
        src = dst = None
        if rank == 0:
            src = torch.empty(total * bw, dtype=torch.int64, device="cuda")
            dst = torch.empty(total * bw, dtype=torch.int64, device="cuda")
            ctx.fill_uniform(src.data_ptr(), 0, total * 64)
            ctx.sync()

        def compute(inp, out, nb):
            pkg.encrypted_dct_quantize(ctx, plan, inp.data_ptr(), out.data_ptr(), nb)

        def run_ingest():
            return ingest.scatter_compute_gather(dist, rank, world, total, args.ingest_chunk, bw, src, dst,
                                                 lambda n: torch.empty(n, dtype=torch.int64, device="cuda"), compute)

        run_ingest()                                     # warm-up (NCCL connections, buffers)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        i0.record(stream)
        for _ in range(reps):
            run_ingest()
        i1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tm = torch.tensor([i0.elapsed_time(i1)], device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ok = True
        if rank == 0:                                    # the gathered stream equals the local computation, chunk by chunk
            scratch = torch.empty(args.ingest_chunk * bw, dtype=torch.int64, device="cuda")
            for a in range(0, total, args.ingest_chunk):
                nb = min(args.ingest_chunk, total - a)
                pkg.encrypted_dct_quantize(ctx, plan, src[a * bw:].data_ptr(), scratch.data_ptr(), nb)
                torch.cuda.synchronize()
                ok = ok and bool(torch.equal(scratch[:nb * bw], dst[a * bw:(a + nb) * bw]))
            del scratch
            if not ok:
                raise SystemExit("bench.py: the gathered output of the scatter path differs from the local computation")
        ingest_info = {"what": "input stream resident in rank 0's HBM, chunks scattered to the ranks and results gathered back "
                               "over NCCL point-to-point (3-stage pipeline, 2 buffers per direction per rank), blocks/s of the whole job",
                       "value": total * reps / (float(tm.item()) * 1e-3), "unit": UNIT, "blocks": total, "blocks_per_rank": args.ingest_blocks,
                       "chunk_blocks": args.ingest_chunk, "chunk_bytes": args.ingest_chunk * bw * 8, "ms": float(tm.item()) / reps,
                       "bytes_over_nvlink_per_direction": (total - args.ingest_blocks) * bw * 8, "gathered_equals_local": ok}
        del src, dst
        torch.cuda.empty_cache()

    # ---- end to end through the C-ABI with HOST (pinned) buffers, copies inside the call -------
    EB = args.e2e_blocks
    ebytes = EB * ctx.block_words() * 8
    h_in, h_out = pkg.host_alloc(ebytes), pkg.host_alloc(ebytes)
    tmp = torch.empty(EB * ctx.block_words(), dtype=torch.int64, device="cuda")
    ctx.fill_uniform(tmp.data_ptr(), first_block * 64, EB * 64)
    import ctypes as C
    pkg._check(pkg.lib().fhe_download(ctx.h, C.c_void_p(h_in), C.c_void_p(tmp.data_ptr()), ebytes))
    del tmp
    for _ in range(2):
        pkg.encrypted_dct_quantize_host(ctx, plan, h_in, h_out, EB)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 6))
    for _ in range(e2e_steps):
        pkg.encrypted_dct_quantize_host(ctx, plan, h_in, h_out, EB)
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * EB * e2e_steps / float(te.item())
    pkg.host_free(h_in); pkg.host_free(h_out)

    # ---- roofline of the dominant kernel (each ciphertext read once + written once)
    peak, peak_src = measured_peaks()
    ct_bytes = ctx.ct_bytes(2)
    launch_bytes = 2 * 64 * ct_bytes * B          # a kernel of the path reads and writes the chunk once
    names = ["ntt_fwd", "dct_pointwise", "ntt_inv"]
    dom = max(range(3), key=lambda i: stage[i])
    achieved = launch_bytes / (stage[dom] * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            # dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture, stored per block
            traffic = json.load(f)["per_block_bytes"][names[dom]] * B
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": launch_bytes,
                "stage_ms": dict(zip(names, stage)),
                "path_achieved": value / world * 2 * 64 * ct_bytes / 1e9,
                "path_frac": value / world * 2 * 64 * ct_bytes / 1e9 / peak}
    # what actually bounds the path: constant multiplications `mulc` on the FMA-heavy
    # pipe.  Per block: 2 transforms x 256 limb-polynomials x n/2 log2 n butterflies + 256 per
    # (poly, limb, position) in the pointwise network; peak = tools/ubench (profiles/r01d_ubench.txt).
    logn = N_RING.bit_length() - 1
    mulc_per_block = 2 * 64 * 2 * ctx.k * (N_RING // 2) * logn + 256 * 2 * ctx.k * N_RING
    roofline["int_multiply_pipe"] = {"achieved": value / world * mulc_per_block / 1e9, "peak": 1227.2,
                                     "unit": "G mulc/s per GPU", "frac": value / world * mulc_per_block / 1e9 / 1227.2,
                                     "mulc_per_block": mulc_per_block,
                                     "peak_source": "measured, tools/ubench at 1965 MHz (profiles/r01d_ubench.txt)"}

    cpu, parity = None, None
    if rank == 0 and not args.no_cpu:
        threads = os.cpu_count() or 1
        first = []
        v, secs, done = cpu_sample(threads, 8 * threads, min_seconds=12.0, keep_first=first)
        v1, _, _ = cpu_sample(1, 4)
        # the oracle as the checker: the blocks it just computed are blocks 0.. of rank 0's chunk
        nchk = min(SPOT, 8 * threads)
        if not (first[0][: nchk * ctx.block_words()] == spot[: nchk * ctx.block_words()]).all():
            raise SystemExit("bench.py: GPU output differs from the oracle on the spot-checked blocks")
        parity = {"blocks": nchk, "bit_exact": True, "against": "oracle, blocks 0.. of rank 0's timed chunk"}
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{done} blocks of the same workload in {secs:.1f} s on {threads} threads "
                         f"(oracle, op-by-op reference graph); single thread: {v1:.2f} blocks/s",
               "single_thread_value": v1}

    # ---- the other two hot-path configurations, full size, fewer steps (their own lines: --workload resize /
    # decode); every rank runs them, like the headline
    side = {}
    plan.close()
    ctx.close()
    if not args.no_side:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_workloads
        for key, fn in (("config3_resize", bench_workloads.run_resize), ("config5_decode", bench_workloads.run_decode)):
            try:
                side[key] = fn(args, rank, world, local_rank, torch, dist, stream, ClockSampler, measured_peaks(), side=True)
            except SystemExit as e:
                side[key] = {"error": str(e)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "blocks_per_step_per_gpu": B, "image_blocks": IMAGE_BLOCKS,
                       "l2": "inputs larger than L2 (no flush needed)", "parallelism": f"blocks sharded x{world}",
                       "internal_chunk_blocks": args.chunk},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ebytes, "d2h_bytes_per_step": ebytes,
                    "blocks_per_step": EB, "steps": e2e_steps},
            "roofline": roofline, "cpu_baseline": cpu, "server_jpeg_tile": tile_info,
            "output_checksum_xor": hex(int(cs_t.item())), "parity_spot_check": parity, "ingest_scatter": ingest_info, **side,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
