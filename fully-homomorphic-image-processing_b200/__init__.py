"""fully-homomorphic-image-processing_b200 -- Python host mirror over the C-ABI of libfhe_b200.so.

The product is the CUDA library (csrc/, built in-tree as libfhe_b200.so) behind the C-ABI in
include/fhe_b200.h; the reference-facing host layer is C++ (host/seal_shim.hpp,
host/fhe_image_b200.hpp).  This module is the thin ctypes mirror used by tests/ and bench.py so
that the parity tests call through the same C-ABI a reference maintainer would bind.  Names
follow the reference: `encrypted_dct`, `quantize_fhe` (homo/fhe_image.h:196,294),
`FractionalEncoder` (server_jpeg.cpp:100), Evaluator ops (add/sub/multiply_plain ...).

There is NO CPU fallback: if the library is missing or no CUDA device is usable every entry
point raises.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FHE_B200_LIB") or os.path.join(_HERE, "libfhe_b200.so")  # override: A/B builds of the same sources
ROOT = os.path.dirname(_HERE)

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# fhe_image.h:221-236 -- the 12 distinct constants of encrypted_dct, in plan order
DCT_CONSTANTS = (0.541196100, 0.765366865, -1.847759065, 1.175875602, 0.298631336, 2.053119869,
                 3.072711026, 1.501321110, -0.899976223, -2.562915447, -1.961570560, -0.390180644)
# fhe_image.h:99-100
YQT = (16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
       14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113,
       92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99)
UVQT = (17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
        47, 66, 99, 99, 99, 99, 99, 99) + (99,) * 32
# fhe_image.h:22-26
POLY_BASE, N_FRACTIONAL_COEFFS, N_NUMBER_COEFFS, PLAIN_MODULUS = 2, 100, 100, 1 << 14


class FheError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a into the in-tree libfhe_b200.so (nvcc cross-compiles without a GPU)."""
    if os.environ.get("FHE_B200_LIB"):   # an A/B build supplied by the caller: never rebuilt from here
        return LIB_PATH
    src = os.path.join(_HERE, "csrc", "fhe_b200.cu")
    deps = [src] + [os.path.join(_HERE, "csrc", f) for f in (
        "ntt.cuh", "ntt_generic.cuh", "dct.cuh", "modarith.cuh", "ring_host.hpp", "behz.cuh", "behz_host.hpp", "behz_api.inc",
        "resize_api.inc", "jpeg_api.inc")]
    deps.append(os.path.join(ROOT, "include", "fhe_b200.h"))
    stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        subprocess.check_call(["nvcc", *NVCC_FLAGS, "-o", LIB_PATH, src])
    return LIB_PATH


_lib = None
u64p = C.POINTER(C.c_uint64)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FheError(-2, f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    sig = {
        "fhe_last_error": (C.c_char_p, []),
        "fhe_version": (i32, []),
        "fhe_ctx_create": (i32, [u32, u32, u64p, u64, i32, vp, C.POINTER(vp)]),
        "fhe_ctx_destroy": (None, [vp]),
        "fhe_ctx_sync": (i32, [vp]),
        "fhe_default_coeff_modulus": (i32, [u32, u64p, C.POINTER(u32)]),
        "fhe_ct_bytes": (C.c_size_t, [vp, u32]),
        "fhe_device_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "fhe_device_free": (i32, [vp, vp]),
        "fhe_host_alloc": (i32, [C.c_size_t, C.POINTER(vp)]),
        "fhe_host_free": (i32, [vp]),
        "fhe_upload": (i32, [vp, vp, vp, C.c_size_t]),
        "fhe_download": (i32, [vp, vp, vp, C.c_size_t]),
        "fhe_device_copy": (i32, [vp, vp, vp, C.c_size_t]),
        "fhe_device_copy_2d": (i32, [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t]),
        "fhe_frac_encode": (i32, [u32, u64, C.c_double, u32, u32, u32, u64p]),
        "fhe_add": (i32, [vp, vp, vp, u64, u32]),
        "fhe_sub": (i32, [vp, vp, vp, u64, u32]),
        "fhe_negate": (i32, [vp, vp, u64, u32]),
        "fhe_ptab_create": (i32, [vp, u64p, u32, C.POINTER(vp)]),
        "fhe_ptab_destroy": (None, [vp]),
        "fhe_multiply_plain": (i32, [vp, vp, u64, u32, vp, u32]),
        "fhe_add_plain": (i32, [vp, vp, u64, u32, vp, u32]),
        "fhe_sub_plain": (i32, [vp, vp, u64, u32, vp, u32]),
        "fhe_ntt_forward": (i32, [vp, vp, u64]),
        "fhe_ntt_inverse": (i32, [vp, vp, u64]),
        "fhe_dct_plan_create": (i32, [vp, u64p, u64p, u64p, C.POINTER(vp)]),
        "fhe_dct_plan_destroy": (None, [vp]),
        "fhe_dct8x8_quantize": (i32, [vp, vp, vp, vp, u64]),
        "fhe_dct8x8_quantize_host": (i32, [vp, vp, vp, vp, u64]),
        "fhe_fill_uniform": (i32, [vp, vp, u64, u64, u32, u64]),
        "fhe_checksum": (i32, [vp, vp, u64, u64p]),
        "fhe_launch_count": (u64, [vp]),
        "fhe_set_dct_chunk_blocks": (i32, [vp, u32]),
        "fhe_set_stage_timing": (i32, [vp, i32]),
        "fhe_get_stage_times": (i32, [vp, C.POINTER(C.c_float)]),
        "fhe_jpeg_plan_create": (i32, [vp, u64p, u64p, u64p, u64p, u64p, u64p, C.POINTER(vp)]),
        "fhe_jpeg_plan_destroy": (None, [vp]),
        "fhe_jpeg_rgb_tiles": (i32, [vp, vp, vp, vp, u64]),
        "fhe_device_zero": (i32, [vp, vp, C.c_size_t]),
        "fhe_encrypt_transparent": (i32, [vp, u64p, u32, vp]),
        "fhe_ntt_loge": (u32, [u32]),
        "fhe_multiply": (i32, [vp, vp, u32, vp, u32, vp, u64]),
        "fhe_square": (i32, [vp, vp, u32, vp, u64]),
        "fhe_form_create": (i32, [vp, vp, u32, u64, C.POINTER(vp)]),
        "fhe_form_destroy": (None, [vp]),
        "fhe_multiply_form": (i32, [vp, vp, u32, vp, u32, vp, u64]),
        "fhe_upload_async": (i32, [vp, vp, vp, C.c_size_t]),
        "fhe_download_async": (i32, [vp, vp, vp, C.c_size_t]),
        "fhe_transfer_fence": (i32, [vp, u32]),
        "fhe_upload_fence": (i32, [vp, u32]),
        "fhe_device_broadcast": (i32, [vp, vp, vp, u32, u64]),
        "fhe_add_plain_each": (i32, [vp, vp, u64, u32, u64p, u32]),
        "fhe_profile_begin": (i32, [vp]),
        "fhe_profile_end": (i32, [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_float), u64p, u32, C.POINTER(u32)]),
        "fhe_resize_plan_create": (i32, [vp, u64p, C.POINTER(vp)]),
        "fhe_resize_plan_destroy": (None, [vp]),
        "fhe_cubic": (i32, [vp, vp, C.POINTER(vp), u32, vp, vp, vp, u32, vp, u64]),
        "fhe_linear": (i32, [vp, C.POINTER(vp), u32, vp, vp, u32, vp, u64]),
        "fhe_evk_digits": (u32, [vp, u32]),
        "fhe_evk_create": (i32, [vp, u64p, u32, u32, C.POINTER(vp)]),
        "fhe_evk_destroy": (None, [vp]),
        "fhe_relinearize": (i32, [vp, vp, u32, vp, vp, u64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "fhe_last_error", "fhe_version", "fhe_ctx_create", "fhe_ctx_destroy", "fhe_ctx_sync",
    "fhe_default_coeff_modulus", "fhe_ct_bytes", "fhe_device_alloc", "fhe_device_free", "fhe_host_alloc",
    "fhe_host_free", "fhe_upload", "fhe_download", "fhe_device_copy", "fhe_device_copy_2d", "fhe_frac_encode", "fhe_add", "fhe_sub", "fhe_negate",
    "fhe_ptab_create", "fhe_ptab_destroy", "fhe_multiply_plain", "fhe_add_plain", "fhe_sub_plain",
    "fhe_ntt_forward", "fhe_ntt_inverse", "fhe_dct_plan_create", "fhe_dct_plan_destroy",
    "fhe_dct8x8_quantize", "fhe_dct8x8_quantize_host", "fhe_fill_uniform", "fhe_checksum",
    "fhe_launch_count", "fhe_set_dct_chunk_blocks", "fhe_set_stage_timing", "fhe_get_stage_times",
    "fhe_jpeg_plan_create", "fhe_jpeg_plan_destroy", "fhe_jpeg_rgb_tiles", "fhe_device_zero", "fhe_encrypt_transparent", "fhe_ntt_loge", "fhe_multiply", "fhe_square", "fhe_form_create", "fhe_form_destroy", "fhe_multiply_form", "fhe_profile_begin", "fhe_profile_end", "fhe_device_broadcast", "fhe_add_plain_each", "fhe_upload_async", "fhe_download_async", "fhe_transfer_fence", "fhe_upload_fence", "fhe_resize_plan_create", "fhe_resize_plan_destroy", "fhe_cubic", "fhe_linear", "fhe_evk_digits", "fhe_evk_create", "fhe_evk_destroy", "fhe_relinearize",
)


def _check(rc):
    if rc != 0:
        raise FheError(rc, lib().fhe_last_error().decode())


def profile_end_handle(h) -> dict:
    names = C.create_string_buffer(4096)
    ms, counts, n = (C.c_float * 64)(), (C.c_uint64 * 64)(), C.c_uint32(0)
    _check(lib().fhe_profile_end(h, names, 4096, ms, counts, 64, C.byref(n)))
    keys = names.value.decode().split("\n")
    return {keys[i]: (float(ms[i]), int(counts[i])) for i in range(n.value)}


def coeff_modulus_128(n: int) -> list[int]:
    """seal::coeff_modulus_128(n) (server_jpeg.cpp:78)."""
    buf, k = (C.c_uint64 * 16)(), C.c_uint32(0)
    _check(lib().fhe_default_coeff_modulus(n, buf, C.byref(k)))
    return [int(buf[i]) for i in range(k.value)]


def _np_ptr(a):
    import numpy as np
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


class Context:
    """EncryptionParameters + SEALContext + Evaluator (server_jpeg.cpp:74-93) on one B200."""

    def __init__(self, n: int, q: list[int] | None = None, t: int = PLAIN_MODULUS, device: int = 0,
                 stream: int | None = None):
        self.n, self.t = n, t
        self.q = list(q) if q is not None else coeff_modulus_128(n)
        self.k = len(self.q)
        arr = (C.c_uint64 * self.k)(*self.q)
        h = C.c_void_p()
        _check(lib().fhe_ctx_create(n, self.k, arr, t, device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().fhe_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- sizes
    def ct_words(self, size=2): return size * self.k * self.n
    def ct_bytes(self, size=2): return 8 * self.ct_words(size)
    def block_words(self): return 64 * self.ct_words(2)

    # ---- buffers
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _check(lib().fhe_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, dptr: int): _check(lib().fhe_device_free(self.h, C.c_void_p(dptr)))

    def upload(self, dptr: int, host):
        _check(lib().fhe_upload(self.h, C.c_void_p(dptr), host.ctypes.data_as(C.c_void_p), host.nbytes))

    def download(self, host, dptr: int):
        _check(lib().fhe_download(self.h, host.ctypes.data_as(C.c_void_p), C.c_void_p(dptr), host.nbytes))
        return host

    def to_device(self, host) -> int:
        d = self.alloc(host.nbytes)
        self.upload(d, host)
        return d

    def sync(self): _check(lib().fhe_ctx_sync(self.h))

    def profile_begin(self): _check(lib().fhe_profile_begin(self.h))

    def profile_end(self) -> dict:
        """{kernel: (summed device ms, launches)} since profile_begin (include/fhe_b200.h)"""
        return profile_end_handle(self.h)
    def launch_count(self) -> int: return int(lib().fhe_launch_count(self.h))
    def set_dct_chunk_blocks(self, blocks: int): _check(lib().fhe_set_dct_chunk_blocks(self.h, blocks))
    def set_stage_timing(self, on: bool): _check(lib().fhe_set_stage_timing(self.h, int(on)))

    def stage_times_ms(self) -> tuple[float, float, float]:
        """(forward NTT, pointwise DCT+quantize, inverse NTT) device ms of the last DCT call."""
        buf = (C.c_float * 3)()
        _check(lib().fhe_get_stage_times(self.h, buf))
        return float(buf[0]), float(buf[1]), float(buf[2])

    # ---- Evaluator (device arrays of n_ct ciphertexts)
    def add(self, a, b, n_ct, size=2): _check(lib().fhe_add(self.h, C.c_void_p(a), C.c_void_p(b), n_ct, size))
    def sub(self, a, b, n_ct, size=2): _check(lib().fhe_sub(self.h, C.c_void_p(a), C.c_void_p(b), n_ct, size))
    def negate(self, a, n_ct, size=2): _check(lib().fhe_negate(self.h, C.c_void_p(a), n_ct, size))

    def multiply_plain(self, a, n_ct, tab: "PlainTable", index=0, size=2):
        _check(lib().fhe_multiply_plain(self.h, C.c_void_p(a), n_ct, size, tab.h, index))

    def add_plain(self, a, n_ct, tab: "PlainTable", index=0, size=2):
        _check(lib().fhe_add_plain(self.h, C.c_void_p(a), n_ct, size, tab.h, index))

    def sub_plain(self, a, n_ct, tab: "PlainTable", index=0, size=2):
        _check(lib().fhe_sub_plain(self.h, C.c_void_p(a), n_ct, size, tab.h, index))

    def multiply(self, a, size_a, b, size_b, dest, n_ct=1):
        """Evaluator::multiply: dest [n_ct][size_a + size_b - 1][k][n] (device pointers)."""
        _check(lib().fhe_multiply(self.h, C.c_void_p(a), size_a, C.c_void_p(b), size_b, C.c_void_p(dest), n_ct))

    def square(self, a, size_a, dest, n_ct=1):
        _check(lib().fhe_square(self.h, C.c_void_p(a), size_a, C.c_void_p(dest), n_ct))

    def form(self, a, size, n_ct=1) -> "OperandForm":
        """the operand half of Evaluator::multiply, done once: BEHZ extension + forward NTTs (fhe_form_create)"""
        return OperandForm(self, a, size, n_ct)

    def multiply_form(self, fa: "OperandForm", a_div, fb: "OperandForm", b_div, dest, n_ct):
        """product i = ciphertext i // a_div of fa times ciphertext i // b_div of fb -> dest (fhe_multiply_form)"""
        _check(lib().fhe_multiply_form(self.h, fa.h, a_div, fb.h, b_div, C.c_void_p(dest), n_ct))

    def cubic(self, plan: "ResizePlan", taps, size_in, ft, ft2, ft3, t_div, dest, n_ct):
        """fused Cubic (fhe_resize.h:143-189) of n_ct calls; taps: n_ct * 4 device pointers (A B C D per call)"""
        arr = (C.c_void_p * (4 * n_ct))(*[int(p) for p in taps])
        _check(lib().fhe_cubic(self.h, plan.h, arr, size_in, ft.h, ft2.h, ft3.h, t_div, C.c_void_p(dest), n_ct))

    def linear(self, taps, size_in, f1mt, ft, t_div, dest, n_ct):
        """fused Linear (fhe_resize.h:191-204) of n_ct calls; taps: n_ct * 2 device pointers (A B per call)"""
        arr = (C.c_void_p * (2 * n_ct))(*[int(p) for p in taps])
        _check(lib().fhe_linear(self.h, arr, size_in, f1mt.h, ft.h, t_div, C.c_void_p(dest), n_ct))

    def evk_digits(self, dbc: int) -> int: return int(lib().fhe_evk_digits(self.h, dbc))

    def relinearize(self, a, size, evk: "EvaluationKeys", dest, n_ct=1):
        _check(lib().fhe_relinearize(self.h, C.c_void_p(a), size, evk.h, C.c_void_p(dest), n_ct))

    def ntt_forward(self, a, n_limbpolys): _check(lib().fhe_ntt_forward(self.h, C.c_void_p(a), n_limbpolys))
    def ntt_inverse(self, a, n_limbpolys): _check(lib().fhe_ntt_inverse(self.h, C.c_void_p(a), n_limbpolys))

    def encrypt_transparent(self, plains, dptr):
        """(Delta * m, 0) for every row of `plains` ([count][n], coefficients < t) into dptr."""
        import numpy as np
        arr = np.ascontiguousarray(plains, dtype=np.uint64).reshape(-1, self.n)
        _check(lib().fhe_encrypt_transparent(self.h, _np_ptr(arr), arr.shape[0], C.c_void_p(dptr)))

    # ---- synthetic inputs / checks
    def fill_uniform(self, dptr, first_ct, n_ct, size=2, seed=0x5EA1B200):
        _check(lib().fhe_fill_uniform(self.h, C.c_void_p(dptr), first_ct, n_ct, size, seed))

    def checksum(self, dptr, count) -> int:
        out = C.c_uint64(0)
        _check(lib().fhe_checksum(self.h, C.c_void_p(dptr), count, C.byref(out)))
        return int(out.value)


class FractionalEncoder:
    """seal::FractionalEncoder(plain_modulus, poly_modulus, int_coeffs, frac_coeffs, base), server_jpeg.cpp:100."""

    def __init__(self, n: int, t: int = PLAIN_MODULUS, int_coeffs=N_NUMBER_COEFFS,
                 frac_coeffs=N_FRACTIONAL_COEFFS, base=POLY_BASE):
        self.n, self.t, self.ic, self.fc, self.base = n, t, int_coeffs, frac_coeffs, base

    def encode(self, value: float):
        import numpy as np
        out = np.zeros(self.n, dtype=np.uint64)
        _check(lib().fhe_frac_encode(self.n, self.t, float(value), self.ic, self.fc, self.base, _np_ptr(out)))
        return out


class PlainTable:
    """A set of plaintexts pre-transformed for multiply_plain / add_plain / sub_plain."""

    def __init__(self, ctx: Context, plains):
        import numpy as np
        arr = np.ascontiguousarray(plains, dtype=np.uint64).reshape(-1, ctx.n)
        h = C.c_void_p()
        _check(lib().fhe_ptab_create(ctx.h, _np_ptr(arr), arr.shape[0], C.byref(h)))
        self.h, self.ctx, self.count = h, ctx, arr.shape[0]

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().fhe_ptab_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ResizePlan:
    """fhe_resize_plan: the five plaintexts of Cubic from the caller's encoder (fhe_resize.h:151-183)."""

    def __init__(self, ctx: "Context", enc: "FractionalEncoder"):
        import numpy as np
        plains = np.ascontiguousarray(np.stack([enc.encode(v) for v in (3.0, 2.0, 5.0, 4.0, 0.5)]), dtype=np.uint64)
        self.h = C.c_void_p()
        _check(lib().fhe_resize_plan_create(ctx.h, plains.ctypes.data_as(u64p), C.byref(self.h)))

    def close(self):
        if self.h:
            lib().fhe_resize_plan_destroy(self.h)
            self.h = C.c_void_p()


class OperandForm:
    """fhe_form: n_ct ciphertexts of `size` polynomials in the NTT domain of q and B_sk (include/fhe_b200.h)."""

    def __init__(self, ctx: "Context", a, size: int, n_ct: int = 1):
        self.h = C.c_void_p()
        self.size, self.n_ct = size, n_ct
        _check(lib().fhe_form_create(ctx.h, C.c_void_p(a), size, n_ct, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().fhe_form_destroy(self.h)
            self.h = C.c_void_p()


class EvaluationKeys:
    """seal::EvaluationKeys of generate_evaluation_keys(dbc, count) (tests/parameters.cpp:89-90),
    given in coefficient form [count][k][digits][2][k][n] and kept NTT-transformed on the device."""

    def __init__(self, ctx: Context, evk, dbc: int, count: int):
        import numpy as np
        arr = np.ascontiguousarray(evk, dtype=np.uint64).reshape(-1)
        assert arr.size == count * ctx.k * ctx.evk_digits(dbc) * 2 * ctx.k * ctx.n
        h = C.c_void_p()
        _check(lib().fhe_evk_create(ctx.h, _np_ptr(arr), dbc, count, C.byref(h)))
        self.h, self.ctx, self.dbc, self.count = h, ctx, dbc, count

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().fhe_evk_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DctPlan:
    """Tables of encrypted_dct (+ quantize_fhe when `quant` is given), fhe_image.h:196-305."""

    def __init__(self, ctx: Context, encoder: FractionalEncoder, quant=None):
        import numpy as np
        consts = np.stack([encoder.encode(c) for c in DCT_CONSTANTS])
        eighth = encoder.encode(0.125)
        qp = None
        if quant is not None:
            assert len(quant) == 64
            qp = np.stack([encoder.encode(1 / float(v)) for v in quant])  # fhe_image.h:301
        h = C.c_void_p()
        _check(lib().fhe_dct_plan_create(ctx.h, _np_ptr(consts), _np_ptr(eighth),
                                         _np_ptr(qp) if qp is not None else None, C.byref(h)))
        self.h, self.ctx = h, ctx

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().fhe_dct_plan_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# fhe_image.h:317-319 -- the plaintexts rgb_to_ycc_fhe encodes, in call order
COLOR_CONSTANTS = (0.299, 0.587, 0.114, -0.168736, 0.331264, 0.5, 0.5, 0.418688, 0.081312)


class JpegPlan:
    """Tables of the whole server_jpeg tile step (server_jpeg.cpp:126-135): rgb_to_ycc_fhe +
    encrypted_dct per channel (+ quantize_fhe with `quant_y` / `quant_c` when given)."""

    def __init__(self, ctx: Context, encoder: FractionalEncoder, quant_y=None, quant_c=None):
        import numpy as np
        color = np.stack([encoder.encode(c) for c in COLOR_CONSTANTS])
        off = encoder.encode(128.0)
        consts = np.stack([encoder.encode(c) for c in DCT_CONSTANTS])
        eighth = encoder.encode(0.125)
        qy = np.stack([encoder.encode(1 / float(v)) for v in quant_y]) if quant_y is not None else None
        qc = np.stack([encoder.encode(1 / float(v)) for v in quant_c]) if quant_c is not None else None
        h = C.c_void_p()
        _check(lib().fhe_jpeg_plan_create(ctx.h, _np_ptr(color), _np_ptr(off), _np_ptr(consts), _np_ptr(eighth),
                                          _np_ptr(qy) if qy is not None else None,
                                          _np_ptr(qc) if qc is not None else None, C.byref(h)))
        self.h, self.ctx = h, ctx

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().fhe_jpeg_plan_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def jpeg_rgb_tiles(ctx: Context, plan: JpegPlan, d_in: int, d_out: int, ntiles: int):
    """rgb_to_ycc_fhe + encrypted_dct (+ quantize_fhe) on device tiles [tile][3][64][2][k][n]."""
    _check(lib().fhe_jpeg_rgb_tiles(ctx.h, plan.h, C.c_void_p(d_in), C.c_void_p(d_out), ntiles))


def encrypted_dct_quantize(ctx: Context, plan: DctPlan, d_in: int, d_out: int, nblocks: int):
    """encrypted_dct (+ quantize_fhe) on device buffers [block][64][2][k][n]; d_in may equal d_out."""
    _check(lib().fhe_dct8x8_quantize(ctx.h, plan.h, C.c_void_p(d_in), C.c_void_p(d_out), nblocks))


def encrypted_dct_quantize_host(ctx: Context, plan: DctPlan, h_in, h_out, nblocks: int):
    """Same through HOST buffers (numpy arrays or raw pointers); copies are inside the call."""
    pin = h_in if isinstance(h_in, int) else h_in.ctypes.data
    pout = h_out if isinstance(h_out, int) else h_out.ctypes.data
    _check(lib().fhe_dct8x8_quantize_host(ctx.h, plan.h, C.c_void_p(pin), C.c_void_p(pout), nblocks))


def host_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    _check(lib().fhe_host_alloc(nbytes, C.byref(p)))
    return p.value


def host_free(ptr: int): _check(lib().fhe_host_free(C.c_void_p(ptr)))


# ---- multi-GPU: blocks are independent, so ranks take contiguous block ranges ----
def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """[first, last) units of `rank` out of `world` (rank r gets [r*N/G, (r+1)*N/G))."""
    if world < 1 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard arguments")
    return (rank * n_units) // world, ((rank + 1) * n_units) // world
