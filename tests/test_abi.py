"""C-ABI surface: the library loads, exports every symbol include/fhe_b200.h declares, the host
logic (encoder, default chain, argument checks) works, and compute entry points FAIL LOUDLY
without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, PKG_NAME


def header_functions():
    src = open(os.path.join(ROOT, "include", "fhe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fhe_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol(pkg):
    path = pkg.build()
    L = C.CDLL(path)
    names = header_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/fhe_b200.h but not exported"
    assert sorted(pkg.EXPORTED_SYMBOLS) == names


def test_library_is_sm100a_only(pkg):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", pkg.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_touches_the_oracle():
    pk = os.path.join(ROOT, PKG_NAME)
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                for line in txt.splitlines():
                    code = line.split("//")[0].split("#")[0] if f.endswith(".py") else line.split("//")[0]
                    assert "liboracle" not in code and "/root/reference" not in code, (f, line)
                    if re.match(r"\s*(#\s*include|import|from)\b", code):
                        assert "oracle" not in code and "pymodel" not in code, (f, line)
                    assert "dlopen" not in code, (f, line)


def test_default_chain_and_encoder_host_logic(pkg, oracle_mod):
    assert pkg.coeff_modulus_128(4096) == oracle_mod.default_coeff_modulus(4096)
    assert pkg.coeff_modulus_128(16384) == oracle_mod.default_coeff_modulus(16384)
    assert pkg.coeff_modulus_128(2048) == [0x3fffffff000001]
    with pytest.raises(pkg.FheError):
        pkg.coeff_modulus_128(1000)
    o = oracle_mod.Oracle(4096)
    enc = pkg.FractionalEncoder(4096)
    vals = list(pkg.DCT_CONSTANTS) + [0.125, 128.0, -128.0, 0.299, 0.587, 0.114, -0.168736, 0.331264, 0.5,
                                      0.418688, 0.081312, 3.0, 5.0, 0.0, 2 ** -60, -7.75]
    vals += [1 / v for v in pkg.YQT + pkg.UVQT]
    for v in vals:
        assert (enc.encode(v) == o.encode(v)).all(), v
    with pytest.raises(pkg.FheError):
        pkg.FractionalEncoder(4096, base=3).encode(1.0)
    with pytest.raises(pkg.FheError):
        pkg.FractionalEncoder(64, int_coeffs=40, frac_coeffs=40).encode(1.0)
    with pytest.raises(pkg.FheError):
        pkg.FractionalEncoder(4096, int_coeffs=4).encode(100.0)


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.FheError) as ei:
        pkg.Context(4096)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_invalid_parameters_are_rejected_before_touching_the_device(pkg):
    for n, q in [(1000, [0x7fffffff380001]), (4096, [0x7fffffff380003]), (4096, [0x7fffffff380001] * 2),
                 (4096, [12289]), (4096, [0x1fffffffffe00001 << 1 | 1])]:
        with pytest.raises(pkg.FheError) as ei:
            pkg.Context(n, q)
        assert ei.value.code in (-1, -3), (n, q, ei.value)


def test_shard_range(pkg):
    for n in (0, 1, 7, 64, 49152, 3145728):
        for w in (1, 2, 3, 4, 8):
            r = [pkg.shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pkg.shard_range(10, 2, 2)
