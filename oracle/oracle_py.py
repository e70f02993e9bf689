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
    srcs = [os.path.join(_HERE, f) for f in ("fhe_oracle.c", "fhe_oracle_behz.c", "fhe_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in srcs
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
        # part 2: BEHZ multiply / relinearize / harness / graphs
        L.orc_behz_create.restype = C.c_void_p
        L.orc_behz_create.argtypes = [C.c_void_p]
        L.orc_behz_destroy.argtypes = [C.c_void_p]
        L.orc_behz_aux_count.restype = C.c_uint32
        L.orc_behz_aux_count.argtypes = [C.c_void_p]
        L.orc_behz_bsk.restype = C.c_uint64
        L.orc_behz_bsk.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_behz_extend.argtypes = [C.c_void_p, u64p, u64p]
        L.orc_behz_scale_round.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.orc_multiply.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p, C.c_uint32, u64p]
        L.orc_square.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p]
        L.orc_keygen.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p]
        L.orc_encrypt.argtypes = [C.c_void_p, u64p, u64p, C.c_uint32, C.c_uint64, u64p]
        L.orc_phase.argtypes = [C.c_void_p, u64p, u64p, C.c_uint32, u64p]
        L.orc_evk_digits.restype = C.c_uint32
        L.orc_evk_digits.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_evk_gen.argtypes = [C.c_void_p, u64p, C.c_uint32, C.c_uint32, C.c_uint64, u64p]
        L.orc_relinearize.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p, C.c_uint32, C.c_uint32, u64p]
        L.orc_cubic.restype = C.c_uint32
        L.orc_cubic.argtypes = [C.c_void_p, C.POINTER(u64p), C.POINTER(C.c_uint32), u64p, C.c_uint32,
                                C.c_uint32, C.c_uint32, u64p]
        L.orc_linear.restype = C.c_uint32
        L.orc_linear.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p, C.c_uint32, u64p, C.c_uint32,
                                 C.c_uint32, C.c_uint32, u64p]
        L.orc_homomorphic_sincos.restype = C.c_uint32
        L.orc_homomorphic_sincos.argtypes = [C.c_void_p, u64p, C.c_uint32, u64p, C.c_int, C.c_uint32,
                                             C.c_uint32, u64p]
        L.orc_approximated_step.restype = C.c_uint32
        L.orc_approximated_step.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_int, C.c_int, C.c_double,
                                            C.c_int, u64p, C.c_uint32, C.c_uint32, u64p]
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

        self._behz = None

    def __del__(self):
        if getattr(self, "_behz", None):
            lib().orc_behz_destroy(self._behz)
            self._behz = None
        if getattr(self, "h", None):
            lib().orc_ctx_destroy(self.h)
            self.h = None

    # -- part 2: ciphertext x ciphertext (BEHZ), relinearize, harness, graphs
    @property
    def behz(self):
        if self._behz is None:
            self._behz = lib().orc_behz_create(self.h)
            if not self._behz:
                raise ValueError("orc_behz_create failed")
        return self._behz

    def bsk(self):
        nb = lib().orc_behz_aux_count(self.behz) + 1
        return [int(lib().orc_behz_bsk(self.behz, j)) for j in range(nb)]

    def behz_extend(self, x):
        out = np.empty(len(self.bsk()) * self.n, dtype=np.uint64)
        lib().orc_behz_extend(self.behz, _p(x), _p(out))
        return out

    def behz_scale_round(self, xq, xb):
        out = np.empty(self.k * self.n, dtype=np.uint64)
        lib().orc_behz_scale_round(self.behz, _p(xq), _p(xb), _p(out))
        return out

    def multiply(self, a, sa, b, sb):
        out = np.empty(self.ct_words(sa + sb - 1), dtype=np.uint64)
        lib().orc_multiply(self.behz, _p(a), sa, _p(b), sb, _p(out))
        return out

    def square(self, a, sa):
        out = np.empty(self.ct_words(2 * sa - 1), dtype=np.uint64)
        lib().orc_square(self.behz, _p(a), sa, _p(out))
        return out

    def keygen(self, seed=1):
        sk = np.empty(self.k * self.n, dtype=np.uint64)
        pk = np.empty(2 * self.k * self.n, dtype=np.uint64)
        lib().orc_keygen(self.behz, seed, _p(sk), _p(pk))
        return sk, pk

    def encrypt(self, pk, plain, seed):
        ct = np.empty(self.ct_words(2), dtype=np.uint64)
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        lib().orc_encrypt(self.behz, _p(pk), _p(plain), len(plain), seed, _p(ct))
        return ct

    def phase(self, sk, ct, size):
        out = np.empty(self.k * self.n, dtype=np.uint64)
        lib().orc_phase(self.behz, _p(sk), _p(ct), size, _p(out))
        return out

    def decrypt(self, sk, ct, size, return_budget=False):
        """round(t x / q) mod t on exact integers from the RNS phase."""
        from math import prod
        ph = self.phase(sk, ct, size).reshape(self.k, self.n)
        q = prod(self.q)
        hats = [q // qi for qi in self.q]
        invs = [pow(h % qi, -1, qi) for h, qi in zip(hats, self.q)]
        out, worst = np.empty(self.n, dtype=np.uint64), 0
        for c in range(self.n):
            x = sum(int(ph[i, c]) * invs[i] % self.q[i] * hats[i] for i in range(self.k)) % q
            num = self.t * x
            mval = (2 * num + q) // (2 * q)
            worst = max(worst, abs(num - mval * q))
            out[c] = mval % self.t
        if return_budget:
            import math
            budget = math.log2(q) - 1 - (math.log2(worst) if worst else 0) if worst else float("inf")
            return out, budget
        return out

    def evk_digits(self, dbc): return int(lib().orc_evk_digits(self.behz, dbc))

    def evk_gen(self, sk, dbc, count, seed=7):
        nd = self.evk_digits(dbc)
        evk = np.empty(count * self.k * nd * 2 * self.k * self.n, dtype=np.uint64)
        lib().orc_evk_gen(self.behz, _p(sk), dbc, count, seed, _p(evk))
        return evk

    def relinearize(self, ct, size, evk, dbc, evk_count):
        out = np.empty(self.ct_words(2), dtype=np.uint64)
        lib().orc_relinearize(self.behz, _p(ct), size, _p(evk), dbc, evk_count, _p(out))
        return out

    def cubic(self, A, B, Cc, D, sizes, t, st=2):
        ins = [np.ascontiguousarray(x, dtype=np.uint64) for x in (A, B, Cc, D)]
        ptrs = (u64p * 4)(*[_p(x) for x in ins])
        szs = (C.c_uint32 * 4)(*sizes)
        out = np.empty(self.ct_words(max(sizes) + 2 * st - 1 + 2), dtype=np.uint64)
        rs = lib().orc_cubic(self.behz, ptrs, szs, _p(t), st, self.ic, self.fc, _p(out))
        return out[: self.ct_words(rs)].copy(), rs

    def linear(self, A, sA, B, sB, t, st=2):
        out = np.empty(self.ct_words(max(sA, sB) + st), dtype=np.uint64)
        rs = lib().orc_linear(self.behz, _p(A), sA, _p(B), sB, _p(t), st, self.ic, self.fc, _p(out))
        return out[: self.ct_words(rs)].copy(), rs

    def homomorphic_sincos(self, x, sx, zero, is_cos):
        out = np.empty(self.ct_words(8 * sx), dtype=np.uint64)
        rs = lib().orc_homomorphic_sincos(self.behz, _p(x), sx, _p(zero), int(is_cos), self.ic, self.fc, _p(out))
        return out[: self.ct_words(rs)].copy(), rs

    def approximated_step(self, amplitude, index, count, order, degree, delta, npos, zeros):
        out = np.empty(npos * self.ct_words(22), dtype=np.uint64)
        rs = lib().orc_approximated_step(self.behz, _p(amplitude), _p(index), _p(count), order, degree,
                                         float(delta), npos, _p(zeros), self.ic, self.fc, _p(out))
        return out[: npos * self.ct_words(rs)].copy(), rs

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
