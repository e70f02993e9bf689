"""The C++ host layer (host/seal_b200.hpp + host/fhe_image_b200.hpp): seal::-shaped classes and
the reference's fhe_image.h call surface over the C-ABI.  CPU: it compiles with g++ against the
public header and fails loudly without a device.  GPU: tests/cpp/test_host_shim.cpp passes and
its fused output equals the CPU oracle's op-by-op evaluation bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, PKG_NAME

SRC = os.path.join(ROOT, "tests", "cpp", "test_host_shim.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_shim")


def build_exe(pkg):
    pkg.build()
    libdir = os.path.join(ROOT, PKG_NAME)
    deps = [SRC, os.path.join(libdir, "host", "seal_b200.hpp"), os.path.join(libdir, "host", "fhe_image_b200.hpp"),
            os.path.join(ROOT, "include", "fhe_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", EXE, SRC, "-L" + libdir,
                               "-lfhe_b200", "-Wl,-rpath," + libdir])
    return EXE


def test_host_layer_compiles_and_has_no_cpu_fallback(pkg):
    exe = build_exe(pkg)
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2048, 4096])
def test_host_layer_matches_oracle_on_gpu(pkg, oracle_mod, n, tmp_path):
    exe = build_exe(pkg)
    dump = str(tmp_path / "dump.bin")
    r = subprocess.run([exe, str(n), dump], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("PASS")
    raw = np.fromfile(dump, dtype=np.uint64)
    x, fused = raw[: raw.size // 2].copy(), raw[raw.size // 2:]
    o = oracle_mod.Oracle(n, oracle_mod.default_coeff_modulus(n))
    assert (x == o.fill_uniform(0, 64)).all()
    o.encrypted_dct(x)
    o.quantize_fhe(x, pkg.YQT)
    assert (x == fused).all()
