"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/fhe_oracle.h.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the
product package never imports this module.  PARITY UNPINNED (SEAL absent).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fhe_oracle.c")
    hdr = os.path.join(_HERE, "fhe_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_ctx_create.restype = C.c_void_p
        L.orc_ctx_create.argtypes = [C.c_uint32, C.c_uint32, u64p, C.c_uint64]
        L.orc_ctx_destroy.argtypes = [C.c_void_p]
        for name in ("orc_psi", "orc_delta_mod", "orc_q"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p, C.c_uint32]
        L.orc_q_mod_t.restype = C.c_uint64
        L.orc_q_mod_t.argtypes = [C.c_void_p]
        L.orc_default_coeff_modulus.restype = C.c_uint32
        L.orc_default_coeff_modulus.argtypes = [C.c_uint32, u64p]
        L.orc_ntt_forward.argtypes = [C.c_void_p, C.c_uint32, u64p]
        L.orc_ntt_inverse.argtypes = [C.c_void_p, C.c_uint32, u64p]
        L.orc_negacyclic_mul.argtypes = [C.c_void_p, C.c_uint32, u64p, u64p, u64p]
        for name in ("orc_add", "orc_sub"):
            getattr(L, name).argtypes = [C.c_void_p, u64p, u64p, C.c_uint32]
        L.orc_negate.argtypes = [C.c_void_p, u64p, C.c_uint32]
        for name in ("orc_add_plain", "orc_sub_plain"):
            getattr(L, name).argtypes = [C.c_void_p, u64p, u64p, C.c_uint32]
        L.orc_multiply_plain.restype = C.c_int
        L.orc_multiply_plain.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p, C.c_uint32]
        L.orc_frac_encode.argtypes = [C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, u64p]
        L.orc_int_encode.argtypes = [C.c_void_p, C.c_int64, u64p]
        L.orc_frac_decode.restype = C.c_double
        L.orc_frac_decode.argtypes = [C.c_void_p, u64p]
        L.orc_encrypted_dct.argtypes = [C.c_void_p, u64p, C.c_uint32, C.c_uint32]
        L.orc_quantize_fhe.argtypes = [C.c_void_p, u64p, f64p, C.c_uint32, C.c_uint32]
        L.orc_rgb_to_ycc_fhe.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_uint32, C.c_uint32]
        L.orc_fill_uniform.argtypes = [C.c_void_p, u64p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
        L.orc_checksum.restype = C.c_uint64
        L.orc_checksum.argtypes = [u64p, C.c_uint64]
        L.orc_bench_dct_quant.restype = C.c_double
        L.orc_bench_dct_quant.argtypes = [C.c_void_p, u64p, C.c_uint64, f64p, C.c_uint32, C.c_uint32, C.c_int]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def default_coeff_modulus(n: int) -> list[int]:
    buf = (C.c_uint64 * 16)()
    k = lib().orc_default_coeff_modulus(n, buf)
    if k == 0:
        raise ValueError(f"no default coeff modulus for n={n}")
    return [int(buf[i]) for i in range(k)]


class Oracle:
    """CPU oracle context: (n, {q_j}, t) like SEALContext (server_jpeg.cpp:74-80)."""

    def __init__(self, n: int, q: list[int] | None = None, t: int = 1 << 14,
                 int_coeffs: int = 100, frac_coeffs: int = 100):
        self.n = n
        self.q = list(q) if q is not None else default_coeff_modulus(n)
        self.k = len(self.q)
        self.t = t
        self.ic, self.fc = int_coeffs, frac_coeffs
        arr = (C.c_uint64 * self.k)(*self.q)
        self.h = lib().orc_ctx_create(n, self.k, arr, t)
        if not self.h:
            raise ValueError("orc_ctx_create rejected the parameters")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ctx_destroy(self.h)
            self.h = None

    # -- introspection
    def psi(self, j): return int(lib().orc_psi(self.h, j))
    def delta_mod(self, j): return int(lib().orc_delta_mod(self.h, j))
    def q_mod_t(self): return int(lib().orc_q_mod_t(self.h))

    def ct_words(self, size=2): return size * self.k * self.n

    # -- NTT
    def ntt_forward(self, j, a): lib().orc_ntt_forward(self.h, j, _p(a)); return a
    def ntt_inverse(self, j, a): lib().orc_ntt_inverse(self.h, j, _p(a)); return a

    def negacyclic_mul(self, j, a, b):
        out = np.empty(self.n, dtype=np.uint64)
        lib().orc_negacyclic_mul(self.h, j, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
        return out

    # -- Evaluator subset (in place on a)
    def add(self, a, b, size=2): lib().orc_add(self.h, _p(a), _p(b), size); return a
    def sub(self, a, b, size=2): lib().orc_sub(self.h, _p(a), _p(b), size); return a
    def negate(self, a, size=2): lib().orc_negate(self.h, _p(a), size); return a
    def add_plain(self, a, plain): lib().orc_add_plain(self.h, _p(a), _p(plain), len(plain)); return a
    def sub_plain(self, a, plain): lib().orc_sub_plain(self.h, _p(a), _p(plain), len(plain)); return a

    def multiply_plain(self, a, plain, size=2):
        rc = lib().orc_multiply_plain(self.h, _p(a), size, _p(plain), len(plain))
        if rc != 0:
            raise ValueError("plain cannot be zero")
        return a

    # -- encoder
    def encode(self, value: float) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint64)
        lib().orc_frac_encode(self.h, float(value), self.ic, self.fc, _p(out))
        return out

    def encode_int(self, value: int) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint64)
        lib().orc_int_encode(self.h, int(value), _p(out))
        return out

    def decode(self, plain) -> float:
        return float(lib().orc_frac_decode(self.h, _p(np.ascontiguousarray(plain, dtype=np.uint64))))

    # -- reference graphs
    def encrypted_dct(self, data):
        assert data.size == 64 * self.ct_words()
        lib().orc_encrypted_dct(self.h, _p(data), self.ic, self.fc)
        return data

    def quantize_fhe(self, data, quant):
        qa = np.ascontiguousarray(quant, dtype=np.float64)
        assert qa.size == 64 and data.size == 64 * self.ct_words()
        lib().orc_quantize_fhe(self.h, _p(data), qa.ctypes.data_as(f64p), self.ic, self.fc)
        return data

    def rgb_to_ycc_fhe(self, r, g, b):
        lib().orc_rgb_to_ycc_fhe(self.h, _p(r), _p(g), _p(b), self.ic, self.fc)

    # -- synthetic inputs
    def fill_uniform(self, first_ct, n_ct, size=2, seed=0x5EA1B200):
        out = np.empty(n_ct * self.ct_words(size), dtype=np.uint64)
        lib().orc_fill_uniform(self.h, _p(out), first_ct, n_ct, size, seed)
        return out

    @staticmethod
    def checksum(a) -> int:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
        return int(lib().orc_checksum(_p(a), a.size))

    def bench_dct_quant(self, data, nblocks, quant, nthreads) -> float:
        qa = np.ascontiguousarray(quant, dtype=np.float64)
        return float(lib().orc_bench_dct_quant(self.h, _p(data), nblocks, qa.ctypes.data_as(f64p),
                                               self.ic, self.fc, nthreads))
