"""bench.py's `resize` and `decode` workloads (BASELINE.json configs 3 and 5): the product's C++ host
layer (host/fhe_resize_b200.hpp RowResizer, host/fhe_decode_b200.hpp approximated_step_batched) driven
in-process through tools/libbench_graphs.so, timed with CUDA events on the stream the library launches
on; the oracle (oracle/, CPU) only as the parity checker and as the timed CPU baseline."""
import ctypes as C
import importlib
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "fully-homomorphic-image-processing_b200"
SHIM = os.path.join(ROOT, "tools", "libbench_graphs.so")

RESIZE = {"n": 8192, "ow": 512, "oh": 512, "dw": 256, "dh": 256}
RESIZE_METRIC, RESIZE_UNIT = "encrypted bicubic resize output pixels/sec (RGB: 3 ciphertexts per pixel)", "pixels/s"
RESIZE_WORKLOAD = "server_resize --bicubic 512x512 -> 256x256 RGB, n=8192 k=4 t=2^14, FractionalEncoder(100,100), bands of output rows"
DECODE = {"n": 8192, "npos": 64, "degree": 12, "order": 64, "dbc": 30}
DECODE_METRIC, DECODE_UNIT = "homomorphic run-length step approximations (8x8 block: 64 positions, degree 12)/sec", "pairs/s"
DECODE_WORKLOAD = ("server_decode approximated_step (fhe_decode.h:202-242), 8x8 block = 64 positions, degree 12, order 64, "
                   "n=8192 k=4 t=2^14, every ct x ct product relinearised (dbc 30)")
MULC_PEAK = 1227.2  # G mulc/s per GPU, tools/ubench at 1965 MHz (profiles/r01d_ubench.txt)


def build_shim(force=False):
    """tools/libbench_graphs.so from tools/bench_graphs.cpp (the one recipe: __graft_entry__.build() calls it with
    force=True).  Without force it rebuilds only when a dependency is newer, and a failed rebuild falls back to
    the library that is there (a snapshot of the tree can reorder modification times)."""
    pkg = importlib.import_module(PKG)
    src = os.path.join(ROOT, "tools", "bench_graphs.cpp")
    libdir = os.path.join(ROOT, PKG)
    deps = [src, os.path.join(libdir, "libfhe_b200.so")] + [os.path.join(libdir, "host", h)
                                                          for h in ("seal_b200.hpp", "fhe_resize_b200.hpp", "fhe_decode_b200.hpp")]
    have = os.path.exists(SHIM)
    if not force and have and all(os.path.getmtime(d) <= os.path.getmtime(SHIM) for d in deps):
        return SHIM
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", SHIM + ".tmp", src, "-I" + os.path.join(cuda, "include"),
           "-L" + libdir, "-lfhe_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + libdir]
    try:
        subprocess.check_call(cmd)
        os.replace(SHIM + ".tmp", SHIM)
    except (subprocess.CalledProcessError, OSError) as e:
        if force or not have:
            raise
        print(f"bench: could not rebuild {SHIM} ({e}); using the existing build", file=sys.stderr)
    return SHIM


def shim():
    pkg = importlib.import_module(PKG)
    pkg.build()
    build_shim()
    L = C.CDLL(SHIM)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.bench_graphs_error.restype = C.c_char_p
    L.bres_create.restype = vp
    L.bres_create.argtypes = [C.c_int, vp, C.c_uint32] + [C.c_int] * 9
    L.bres_step.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.bres_spot.argtypes = [vp, C.c_int, C.c_int, C.c_int, u64p, u64p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.bres_ctx.restype = vp
    L.bres_ctx.argtypes = [vp]
    L.bres_destroy.argtypes = [vp]
    L.bdec_create.restype = vp
    L.bdec_create.argtypes = [C.c_int, vp, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32, u64p, C.c_int, C.c_uint64]
    L.bdec_step.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), u64p]
    L.bdec_inputs.argtypes = [vp, u64p, u64p, u64p]
    L.bdec_ctx.restype = vp
    L.bdec_ctx.argtypes = [vp]
    L.bdec_destroy.argtypes = [vp]
    return L


def _ok(L, rc):
    if rc:
        raise SystemExit("bench: " + L.bench_graphs_error().decode())


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    import graphs_py
    return oracle_py, graphs_py


def _timed(torch, dist, world, stream, steps, fn):
    """K steps bracketed by a barrier + synchronize on both sides, CUDA events on the launching stream, MAX over ranks"""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _profile(pkg, ctx_handle, fn):
    pkg._check(pkg.lib().fhe_profile_begin(ctx_handle))
    fn()
    prof = pkg.profile_end_handle(ctx_handle)
    return {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in prof.items() if k}


def _traffic(workload, kernel, launches):
    """measured DRAM bytes of one step's launches of `kernel`: ncu per-launch mean (profiles/traffic.json, captured on
    the default geometry) x the launches the profiler counted in this step; None when there is no capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            per = json.load(f)[workload]["dram_bytes_per_launch"].get(kernel)
        return per * launches if per else None
    except (OSError, KeyError, ValueError):
        return None


def _ntt_mulc(n):
    return (n // 2) * (n.bit_length() - 1)


# ---------------------------------------------------------------------------------------------------------------
# config 3: bicubic resize
# ---------------------------------------------------------------------------------------------------------------
def resize_band_work(rows, dw, k, n, first_src_rows):
    """Cubic calls and constant multiplications (mulc equivalents of the transforms) of one band"""
    L = 2 * k + 1
    hor = first_src_rows * dw * 3            # horizontal Cubics: one per source row of the band x output column x channel
    ver = rows * dw * 3
    ntts = hor * (3 * 2 + 2 * 4 + 3) * L + ver * (3 * 4 + 2 * 6 + 5) * L
    return hor, ver, ntts * _ntt_mulc(n)


def run_resize(args, rank, world, local_rank, torch, dist, stream, sampler_cls, peak, side=False):
    pkg = importlib.import_module(PKG)
    L = shim()
    g = RESIZE
    # side=True: the run bench.py's default line carries under "config3_resize" -- the SAME configuration, fewer steps
    rows = args.rows
    steps, warmup = (2, 3) if side else (args.steps, max(args.warmup, 3))
    y0 = 96 + rank * rows                      # a band away from the clamped borders; each rank its own (weak scaling)
    h = L.bres_create(local_rank, C.c_void_p(stream.cuda_stream), g["n"], g["ow"], g["oh"], g["dw"], g["dh"], 1, 0, y0, y0 + rows, 0)
    if not h:
        raise SystemExit("bench: " + L.bench_graphs_error().decode())
    cs, launches = C.c_ulonglong(0), C.c_ulonglong(0)

    def step():
        _ok(L, L.bres_step(h, C.byref(cs), C.byref(launches)))

    for _ in range(warmup):
        step()
    prof = _profile(pkg, L.bres_ctx(h), step)
    sampler = sampler_cls(local_rank)
    sampler.start()
    ms = _timed(torch, dist, world, stream, steps, step)
    clocks = sampler.stop()
    value = world * rows * g["dw"] * steps / (ms * 1e-3)
    k = 4
    n_src = 2 * rows + 2                       # source rows a band of a 2x reduction touches
    hor, ver, mulc = resize_band_work(rows, g["dw"], k, g["n"], n_src)
    # steady state of the whole image: 2 horizontal rows + 1 vertical row per output row
    _, _, mulc_row = resize_band_work(1, g["dw"], k, g["n"], 2)
    ct = 2 * k * g["n"] * 8
    dom = max(prof, key=lambda kk: prof[kk]["ms"])
    form_poly = (2 * k + 1) * g["n"] * 8
    kernel_bytes = {   # algorithmic bytes of one step's launches of the kernels that can dominate
        "k_tensor_inv": hor * (6 + 11) * form_poly + ver * (12 + 17) * form_poly,      # operand forms read once + products written once
        "k_form_ntt": 2 * (hor * 6 + ver * 12) * form_poly,
        "k_resize_finish": hor * (11 * form_poly + 4 * ct // 2) + ver * (17 * form_poly + 6 * ct // 2),
        "k_resize_prep": hor * (4 * ct + 6 * form_poly) + ver * (4 * 2 * ct + 12 * form_poly),
    }
    achieved = kernel_bytes.get(dom, 0) / (prof[dom]["ms"] * 1e-3) / 1e9 if prof[dom]["ms"] else 0.0
    path_bytes_per_pixel = 3 * 3.5 * (1 << 20)   # 3.5 MiB per channel-pixel with ideal input reuse
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak[0], "unit": "GB/s",
                "frac": achieved / peak[0], "traffic": _traffic("resize", dom, prof[dom]["launches"]) if rows == 8 else None,
                "peak_source": peak[1],
                "algorithmic_bytes_per_step": kernel_bytes.get(dom), "kernel_ms_per_step": prof,
                "path_achieved": value / world * path_bytes_per_pixel / 1e9,
                "path_frac": value / world * path_bytes_per_pixel / 1e9 / peak[0],
                "int_multiply_pipe": {"achieved": mulc * steps / (ms * 1e-3) / 1e9, "peak": MULC_PEAK, "unit": "G mulc/s per GPU",
                                      "frac": mulc * steps / (ms * 1e-3) / 1e9 / MULC_PEAK,
                                      "what": "forward + inverse transforms only (n/2 log2 n butterflies each), per GPU; the BEHZ "
                                              "per-coefficient kernels are extra work on the same pipe",
                                      "transform_floor_pixels_per_s_whole_image": g["dw"] / (mulc_row / (MULC_PEAK * 1e9))}}
    # ---- end to end: source rows from pinned host memory, output rows back to pinned host memory, inside the step
    L.bres_destroy(h)
    # A band's first source rows cannot be hidden behind kernels and neither can its last download; over the
    # whole image (192 output rows in config 3) that fill and drain are ~1 % -- over an 8-row band they were
    # 16 % (profiles/r02_notes.md F).  So the end-to-end leg runs ONE band three times as long as the
    # device-resident step's, its source rows cycling through the pinned rows of an 8-row band.
    e_rows = 3 * rows
    hh = L.bres_create(local_rank, C.c_void_p(stream.cuda_stream), g["n"], g["ow"], g["oh"], g["dw"], g["dh"], 1, 1, y0, y0 + e_rows,
                       2 * rows + 2)
    if not hh:
        raise SystemExit("bench: " + L.bench_graphs_error().decode())
    ecs = C.c_ulonglong(0)

    def estep():
        _ok(L, L.bres_step(hh, C.byref(ecs), None))

    estep()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e_steps = 1
    for _ in range(e_steps):
        estep()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e_rows * g["dw"] * e_steps / float(te.item())
    src_rows_e = 2 * e_rows + 2
    e2e = {"value": e2e_value, "unit": RESIZE_UNIT, "h2d_bytes_per_step": src_rows_e * g["ow"] * 3 * ct,
           "d2h_bytes_per_step": e_rows * g["dw"] * 3 * 3 * ct, "rows_per_step": e_rows, "steps": e_steps,
           "note": "source rows come from pinned host memory (fhe_upload_async, one output row ahead) and every output row goes back "
                   "to pinned host memory (fhe_download_async) inside the step, both beside the kernels"}
    # ---- parity: one output value of the band against the oracle's op-by-op Cubic graph
    parity, cpu = None, None
    if rank == 0 and not args.no_cpu:
        import numpy as np
        oracle_py, graphs_py = _oracle()
        o = oracle_py.Oracle(g["n"])
        W = o.ct_words(2)
        out, taps = np.empty(o.ct_words(6), dtype=np.uint64), np.empty(16 * W, dtype=np.uint64)
        xf, yf = C.c_double(0), C.c_double(0)
        u64p = C.POINTER(C.c_uint64)
        x, y, ch = 37, y0, 1
        _ok(L, L.bres_spot(hh, x, y, ch, out.ctypes.data_as(u64p), taps.ctypes.data_as(u64p), C.byref(xf), C.byref(yf)))

        def enc(v):
            z = np.zeros(W, dtype=np.uint64)
            if v:
                o.add_plain(z, o.encode(v))
            return z

        t0 = time.perf_counter()
        cols = []
        for r in range(4):
            c, rs = o.cubic(*[taps[(4 * r + cc) * W:(4 * r + cc + 1) * W] for cc in range(4)], [2, 2, 2, 2], enc(xf.value))
            cols.append(c)
        want, rs = o.cubic(*cols, [4, 4, 4, 4], enc(yf.value))
        t_cpu = time.perf_counter() - t0
        if not (rs == 6 and (want == out).all()):
            raise SystemExit("bench.py: GPU resize output differs from the oracle on the spot-checked pixel")
        parity = {"what": f"output pixel ({x}, {y}) channel {ch} of the band (size-6 ciphertext)", "bit_exact": True,
                  "against": "oracle op-by-op Cubic graph (5 calls)"}
        # the reference evaluates 15 Cubic calls per output pixel on one thread
        cpu = {"value": 1.0 / (3 * t_cpu), "unit": RESIZE_UNIT, "cores": 1, "kind": "port",
               "sample": f"5 Cubic calls (one channel of one output pixel) in {t_cpu:.2f} s on one thread, the reference's own mode; "
                         f"x3 channels per pixel"}
    L.bres_destroy(hh)
    return {
        "metric": RESIZE_METRIC, "value": value, "unit": RESIZE_UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": RESIZE_WORKLOAD, "output_rows_per_step_per_gpu": rows, "first_output_row": y0,
                   "cubic_calls_per_step": hor + ver, "reference_cubic_calls_for_the_same_rows": rows * g["dw"] * 15,
                   "note": "a band of R rows touches 2R + 2 source rows (the whole image: 2 per output row), so the band rate "
                           "understates the whole-image rate by (3R + 2) / 3R",
                   "l2": "working set of a step (GiB) larger than L2", "parallelism": f"bands of output rows x{world}"},
        "clocks": clocks, "gpu_launches": int(launches.value) * steps, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        "parity_spot_check": parity, "output_checksum": hex(int(cs.value)),
    }


# ---------------------------------------------------------------------------------------------------------------
# config 5: homomorphic run-length decode step
# ---------------------------------------------------------------------------------------------------------------
def decode_keys(n, k, nd, seed=0x5EED):
    """synthetic evaluation keys (timing and bit parity only: any residues below every q_j)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    return rng.integers(0, 1 << 50, size=k * nd * 2 * k * n, dtype=np.uint64)


def run_decode(args, rank, world, local_rank, torch, dist, stream, sampler_cls, peak, side=False):
    import numpy as np
    pkg = importlib.import_module(PKG)
    L = shim()
    g = dict(DECODE)
    steps, warmup = (3, 3) if side else (args.steps, max(args.warmup, 3))   # side: same configuration, fewer steps
    k, nd = 4, 2
    keys = decode_keys(g["n"], k, nd)
    u64p = C.POINTER(C.c_uint64)

    def create(npos, degree, host):
        h = L.bdec_create(local_rank, C.c_void_p(stream.cuda_stream), g["n"], npos, degree, g["order"], g["dbc"],
                          keys.ctypes.data_as(u64p), host, 3 * rank)
        if not h:
            raise SystemExit("bench: " + L.bench_graphs_error().decode())
        return h

    h = create(g["npos"], g["degree"], 0)
    launches, osz = C.c_ulonglong(0), C.c_uint(0)

    def step():
        _ok(L, L.bdec_step(h, C.byref(launches), C.byref(osz), None))

    for _ in range(warmup):
        step()
    prof = _profile(pkg, L.bdec_ctx(h), step)
    sampler = sampler_cls(local_rank)
    sampler.start()
    ms = _timed(torch, dist, world, stream, steps, step)
    clocks = sampler.stop()
    value = world * steps / (ms * 1e-3)
    # work formed per pair: (npos * degree cosine + degree sine) series of 7 products + npos * degree + npos closing products,
    # each product: 2 operand forms at most (x enters 5 products once) and 3 inverse polys, each relinearisation
    # k * nd digit polys transformed in k limbs + 2k inverse transforms
    n_series = g["npos"] * g["degree"] + g["degree"]
    products = 7 * n_series + g["npos"] * g["degree"] + g["npos"]
    Lq = 2 * k + 1
    form_polys = (6 * 2) * n_series + 2 * 2 * g["npos"] * g["degree"] + 2 * g["npos"] + 2     # forms built (size-2 operands)
    ntts = form_polys * Lq + products * 3 * Lq + products * (k * nd * k + 2 * k)
    mulc = ntts * _ntt_mulc(g["n"])
    dom = max(prof, key=lambda kk: prof[kk]["ms"])
    ct = 2 * k * g["n"] * 8
    form_poly = Lq * g["n"] * 8
    kernel_bytes = {
        "k_tensor_inv": products * (4 + 3) * form_poly,
        "k_form_ntt": 2 * form_polys * form_poly,
        "k_relin_digits": products * (ct // 2 // 2 + k * nd * k * g["n"] * 8),
        "k_relin_accum": products * (k * nd * k * g["n"] * 8 + 2 * k * g["n"] * 8),
        "k_behz_scale": products * 3 * (form_poly + k * g["n"] * 8),
    }
    achieved = kernel_bytes.get(dom, 0) / (prof[dom]["ms"] * 1e-3) / 1e9 if prof[dom]["ms"] else 0.0
    path_bytes = (3 + g["npos"]) * ct          # 3 ciphertexts in, npos size-2 ciphertexts out
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak[0], "unit": "GB/s", "frac": achieved / peak[0],
                "traffic": _traffic("decode", dom, prof[dom]["launches"]), "peak_source": peak[1],
                "algorithmic_bytes_per_step": kernel_bytes.get(dom),
                "kernel_ms_per_step": prof, "path_achieved": value / world * path_bytes / 1e9,
                "path_frac": value / world * path_bytes / 1e9 / peak[0],
                "note": "the path's algorithmic HBM bytes are tiny (3 ciphertexts in, 64 out for 10^4 ciphertext products): "
                        "this configuration is bound by the integer-multiply pipe, see int_multiply_pipe",
                "int_multiply_pipe": {"achieved": mulc * steps / (ms * 1e-3) / 1e9, "peak": MULC_PEAK, "unit": "G mulc/s per GPU",
                                      "frac": mulc * steps / (ms * 1e-3) / 1e9 / MULC_PEAK,
                                      "what": "transforms of the products formed (operand forms, inverse transforms, relinearisation "
                                              "digits), n/2 log2 n butterflies each; per-coefficient BEHZ work is extra",
                                      "ct_products_formed_per_pair": products,
                                      "ct_products_in_the_reference_graph_per_pair": 2 * g["npos"] * g["degree"] * 15 + g["npos"] * g["degree"] + g["npos"]}}
    L.bdec_destroy(h)
    # ---- end to end: inputs from pinned host memory, all outputs back to pinned host memory inside the step
    hh = create(g["npos"], g["degree"], 1)

    def estep():
        _ok(L, L.bdec_step(hh, None, None, None))

    estep()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e_steps = max(2, min(steps, 4))
    for _ in range(e_steps):
        estep()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": world * e_steps / float(te.item()), "unit": DECODE_UNIT, "h2d_bytes_per_step": 3 * ct,
           "d2h_bytes_per_step": g["npos"] * ct, "steps": e_steps}
    L.bdec_destroy(hh)
    # ---- parity + CPU baseline: a (1 position, degree 1) step through the same engine against the oracle
    parity, cpu = None, None
    if rank == 0 and not args.no_cpu:
        oracle_py, graphs_py = _oracle()
        o = oracle_py.Oracle(g["n"])
        W = o.ct_words(2)
        hs = create(1, 1, 0)
        amp, idx, cnt, out = (np.empty(W, dtype=np.uint64) for _ in range(4))
        _ok(L, L.bdec_inputs(hs, amp.ctypes.data_as(u64p), idx.ctypes.data_as(u64p), cnt.ctypes.data_as(u64p)))
        _ok(L, L.bdec_step(hs, None, None, out.ctypes.data_as(u64p)))
        L.bdec_destroy(hs)
        zeros = np.zeros(2 * W, dtype=np.uint64)      # transparent encrypt(encode(0.0))
        t0 = time.perf_counter()
        want = graphs_py.approximated_step_relin_oracle(o, amp, idx, cnt, g["order"], 1, 0.5, 1, zeros, keys, g["dbc"])
        t_cpu = time.perf_counter() - t0
        if not (want == out).all():
            raise SystemExit("bench.py: GPU decode output differs from the oracle on the spot-checked step")
        parity = {"what": "approximated_step with 1 position, degree 1 (2 Taylor series, 32 products + relinearisations) through "
                          "the same engine and keys", "bit_exact": True, "against": "oracle op-by-op graph, every product relinearised"}
        full = g["npos"] * g["degree"]
        cpu = {"value": 1.0 / (t_cpu * full), "unit": DECODE_UNIT, "cores": 1, "kind": "port",
               "sample": f"one (position, j) term = 2 Taylor series of 15 products, relinearised, in {t_cpu:.2f} s on one thread "
                         f"(the reference's own mode, oracle primitives composed in Python); a pair has {full} such terms"}
    return {
        "metric": DECODE_METRIC, "value": value, "unit": DECODE_UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": DECODE_WORKLOAD,
                   "positions": g["npos"], "degree": g["degree"], "order": g["order"], "dbc": g["dbc"], "out_ct_size": int(osz.value),
                   "l2": "working set of a step (GiB) larger than L2", "parallelism": f"run-length pairs x{world}"},
        "clocks": clocks, "gpu_launches": int(launches.value) * steps, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        "parity_spot_check": parity,
    }


# ---------------------------------------------------------------------------------------------------------------
# --impl reference: the oracle's op-by-op graphs on the host cores (bounded samples)
# ---------------------------------------------------------------------------------------------------------------
def reference_resize(args):
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    oracle_py, _ = _oracle()
    g = RESIZE
    threads = os.cpu_count() or 1
    o = oracle_py.Oracle(g["n"])
    W = o.ct_words(2)
    data = o.fill_uniform(0, 17, 2, 0xC0F3)
    t = data[16 * W:17 * W]

    def channel_pixel(_):
        cols = [o.cubic(*[data[(4 * r + c) * W:(4 * r + c + 1) * W] for c in range(4)], [2, 2, 2, 2], t)[0] for r in range(4)]
        return o.cubic(*cols, [4, 4, 4, 4], t)[1]

    per_step = threads                          # channel-pixels per step: one per thread (5 Cubic calls each)
    with ThreadPoolExecutor(threads) as ex:
        for _ in range(min(args.warmup, 1)):
            list(ex.map(channel_pixel, range(per_step)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(channel_pixel, range(per_step)))
        secs = time.perf_counter() - t0
    value = per_step * args.steps / 3.0 / secs
    sample = (f"{per_step} channel-pixels/step (5 Cubic calls each: 4 horizontal + 1 vertical, sizes 2 -> 4 -> 6), {threads} threads; "
              f"3 channel-pixels per output pixel")
    return {"impl": "reference", "metric": RESIZE_METRIC, "value": value, "unit": RESIZE_UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": RESIZE_WORKLOAD, "note": "bounded CPU sample of the same workload on the host cores"},
            "cpu_baseline": {"value": value, "unit": RESIZE_UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": RESIZE_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def reference_decode(args):
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    oracle_py, graphs_py = _oracle()
    g = DECODE
    threads = os.cpu_count() or 1
    o = oracle_py.Oracle(g["n"])
    W = o.ct_words(2)
    data = o.fill_uniform(0, 3, 2, 0xDEC)
    keys = decode_keys(g["n"], 4, 2)
    zeros = np.zeros(2 * W, dtype=np.uint64)

    def term(_):
        return graphs_py.approximated_step_relin_oracle(o, data[:W], data[W:2 * W], data[2 * W:], g["order"], 1, 0.5, 1, zeros, keys, g["dbc"])[0]

    per_step = threads                          # (position, j) terms per step: one per thread
    with ThreadPoolExecutor(threads) as ex:
        for _ in range(min(args.warmup, 1)):
            list(ex.map(term, range(per_step)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(term, range(per_step)))
        secs = time.perf_counter() - t0
    full = g["npos"] * g["degree"]
    value = per_step * args.steps / full / secs
    sample = (f"{per_step} (position, j) terms/step (2 Taylor series of 15 products each, every product relinearised), {threads} threads; "
              f"a run-length pair has {full} terms")
    return {"impl": "reference", "metric": DECODE_METRIC, "value": value, "unit": DECODE_UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": DECODE_WORKLOAD, "note": "bounded CPU sample of the same workload on the host cores"},
            "cpu_baseline": {"value": value, "unit": DECODE_UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": DECODE_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
