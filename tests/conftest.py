import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG_NAME = "fully-homomorphic-image-processing_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


def load_pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def emu_lib():
    """CPU emulation of the CUDA thread bodies (tests/emu/emu.cpp), built with g++."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "emu")
    so, src = os.path.join(d, "libfhe_emu.so"), os.path.join(d, "emu.cpp")
    csrc = os.path.join(ROOT, PKG_NAME, "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("ntt.cuh", "ntt_generic.cuh", "dct.cuh", "modarith.cuh", "ring_host.hpp",
                                                      "behz.cuh", "behz_host.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    E = C.CDLL(so)
    u64p = C.POINTER(C.c_uint64)
    E.emu_ring_create.restype = C.c_void_p
    E.emu_ring_create.argtypes = [C.c_uint32, C.c_uint32, u64p, C.c_uint64, C.c_char_p, C.c_int]
    E.emu_ring_destroy.argtypes = [C.c_void_p]
    E.emu_ntt_forward.argtypes = [C.c_void_p, C.c_int, u64p, u64p, C.c_int]
    E.emu_ntt_inverse.argtypes = [C.c_void_p, C.c_int, u64p, u64p, C.c_int]
    E.emu_mulc.restype = C.c_uint64
    E.emu_mulc.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
    for f in ("emu_reduce63", "emu_canon"):
        getattr(E, f).restype = C.c_uint64
        getattr(E, f).argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    E.emu_dct_quant.argtypes = [C.c_void_p, u64p, C.c_uint64, u64p, u64p, u64p]
    for f in ("emu_psi", "emu_delta"):
        getattr(E, f).restype = C.c_uint64
        getattr(E, f).argtypes = [C.c_void_p, C.c_int]
    E.emu_q_mod_t.restype = C.c_uint64
    E.emu_q_mod_t.argtypes = [C.c_void_p]
    return E


def ntt_perm(n):
    """device NTT-domain position of SEAL scrambled index j (csrc/ntt.cuh)."""
    import importlib
    import numpy as np
    loge = int(importlib.import_module(PKG_NAME).lib().fhe_ntt_loge(n))   # the library's own statement of its layout
    assert loge
    e, t = 1 << loge, n >> loge
    j = np.arange(n)
    return (j % e) * t + j // e
