"""The C++ host layer for the resize / decode op libraries (host/fhe_resize_b200.hpp,
host/fhe_decode_b200.hpp): the reference's Cubic / Linear / SampleBicubic / ResizeImage
(homo/fhe_resize.h:143-392) and homomorphic_sin / cos / approximated_step (homo/fhe_decode.h:48-242)
signatures over the C-ABI.  CPU: compiles and fails loudly without a device.  GPU: every result of
tests/cpp/test_resize_decode.cpp equals the oracle's op-by-op graphs bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_resize_decode.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_resize_decode")


def build_exe(pkg):
    pkg.build()
    libdir = os.path.join(ROOT, PKG_NAME)
    deps = [SRC, os.path.join(ROOT, "include", "fhe_b200.h")] + [
        os.path.join(libdir, "host", f) for f in ("seal_b200.hpp", "fhe_resize_b200.hpp", "fhe_decode_b200.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", EXE, SRC, "-L" + libdir,
                               "-lfhe_b200", "-Wl,-rpath," + libdir])
    return EXE


def test_resize_decode_host_layer_compiles_and_has_no_cpu_fallback(pkg):
    exe = build_exe(pkg)
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([exe, "2048", "/dev/null", "/dev/null"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr


def f32(x):
    return np.float32(x)


def resize_bicubic_oracle(o, src, ow, oh, dw, dh, fracs):
    """ResizeImage + SampleBicubic (fhe_resize.h:254-392) restated over the oracle's Cubic, with the
    reference's float arithmetic: u, v are floats, int() truncates toward zero, clamping at the
    borders; the encrypted fractions are inputs (consumed two per output pixel: x then y)."""
    W = o.ct_words(2)

    def px(x, y, ch):
        x = min(max(x, 0), ow - 1)
        y = min(max(y, 0), oh - 1)
        i = (y * ow + x) * 3 + ch
        return src[i * W:(i + 1) * W]

    out, nf = [], 0
    for y in range(dh):
        v = f32(np.float64(f32(y) / f32(dh - 1) * f32(oh)) - 0.5)
        for x in range(dw):
            u = f32(np.float64(f32(x) / f32(dw - 1) * f32(ow)) - 0.5)
            xi, yi = int(u), int(v)          # truncation toward zero, like int(float)
            xf, yf = fracs[nf * W:(nf + 1) * W], fracs[(nf + 1) * W:(nf + 2) * W]
            nf += 2
            for ch in range(3):
                cols = []
                for r in range(4):
                    c, rs = o.cubic(*[px(xi - 1 + cidx, yi - 1 + r, ch) for cidx in range(4)], [2, 2, 2, 2], xf)
                    assert rs == 4
                    cols.append(c)
                res, rs = o.cubic(*cols, [4, 4, 4, 4], yf)
                assert rs == 6
                out.append(res)
    return np.concatenate(out), nf


@pytest.mark.gpu
def test_resize_decode_host_layer_matches_oracle_on_gpu(pkg, oracle_mod, tmp_path):
    exe = build_exe(pkg)
    n = 2048
    o = oracle_mod.Oracle(n, oracle_mod.default_coeff_modulus(n), 1 << 14, 64, 32)
    W = o.ct_words(2)
    n_in = 34 + 90
    data = o.fill_uniform(0, n_in, 2, 0xD1CE)
    # make the five Cubic inputs and the fractions real encryptions, so the result also decrypts
    sk, pk = o.keygen(8)
    vals = [12.0, 200.0, 131.0, 77.0]
    for i, v in enumerate(vals):
        data[i * W:(i + 1) * W] = o.encrypt(pk, o.encode(v), 100 + i)
    data[4 * W:5 * W] = o.encrypt(pk, o.encode(0.625), 200)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    data.tofile(fin)
    r = subprocess.run([exe, str(n), fin, fout], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("PASS")
    got = np.fromfile(fout, dtype=np.uint64)

    def ct(i):
        return data[i * W:(i + 1) * W].copy()

    want = []
    c1, rs = o.cubic(ct(0), ct(1), ct(2), ct(3), [2, 2, 2, 2], ct(4))
    assert rs == 4
    want.append(c1)
    c2, rs = o.cubic(c1, c1, c1, c1, [4, 4, 4, 4], ct(4))
    want.append(c2)
    lin, _ = o.linear(ct(0), 2, ct(1), 2, ct(4))
    want.append(lin)
    s, _ = o.homomorphic_sincos(ct(0), 2, ct(8), 0)
    c, _ = o.homomorphic_sincos(ct(1), 2, ct(9), 1)
    want += [s, c]
    run, rs = o.approximated_step(ct(5), ct(6), ct(7), 8, 2, 0.25, 2, data[8 * W:16 * W].copy())
    assert rs == 22
    want.append(run)
    res, nf = resize_bicubic_oracle(o, data[34 * W:], 5, 6, 3, 3, data[16 * W:34 * W])
    assert nf == 18
    want.append(res)
    # transparent encryption of 0.375: poly 0 = Delta * m, poly 1 = 0
    tz = np.zeros(W, dtype=np.uint64)
    o.add_plain(tz, o.encode(0.375))
    want.append(tz)
    want = np.concatenate(want)
    assert got.size == want.size
    assert (got == want).all()
    # (what the Cubic decrypts to is checked in tests/test_oracle_behz.py with a two-prime modulus;
    # coeff_modulus_128(2048) is one 54-bit prime, too small for two multiplicative levels)
    assert abs(o.decode(o.decrypt(sk, tz, 2)) - 0.375) < 1e-9
