"""The B200-first C++ host layer for the resize / decode servers (host/fhe_resize_b200.hpp: RowResizer
over the fused Cubic / Linear device graphs; host/fhe_decode_b200.hpp: batched Taylor series and step
approximation over operand forms).  CPU: compiles and fails loudly without a device.  GPU: every
result of tests/cpp/test_resize_decode.cpp equals the oracle's op-by-op restatement of the reference's
graphs (homo/fhe_resize.h:143-392, homo/fhe_decode.h:48-242) bit for bit."""
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


from graphs_py import (approximated_step_relin_oracle, cubic_relin_oracle, linear_relin_oracle,  # noqa: E402,F401
                       resize_bicubic_oracle, sin_relin_oracle, sincos_relin_oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("n,ic,fc,dbc", [(2048, 64, 32, 16), (8192, 100, 100, 30)])
def test_resize_decode_host_layer_matches_oracle_on_gpu(pkg, oracle_mod, tmp_path, n, ic, fc, dbc):
    """n = 2048: one limb (k_behz_extend<1,1>); n = 8192: the parameters of BASELINE configs 3 and 5
    (k = 4, FractionalEncoder(100, 100), dbc 30)"""
    exe = build_exe(pkg)
    o = oracle_mod.Oracle(n, oracle_mod.default_coeff_modulus(n), 1 << 14, ic, fc)
    W = o.ct_words(2)
    n_in = 34 + 90
    data = o.fill_uniform(0, n_in, 2, 0xD1CE)
    # make the five Cubic inputs and the fractions real encryptions, so the result also decrypts
    sk, pk = o.keygen(8)
    vals = [12.0, 200.0, 131.0, 77.0]
    for i, v in enumerate(vals):
        data[i * W:(i + 1) * W] = o.encrypt(pk, o.encode(v), 100 + i)
    data[4 * W:5 * W] = o.encrypt(pk, o.encode(0.625), 200)
    fin, fout, fkey = str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(tmp_path / "evk.bin")
    data.tofile(fin)
    evk = o.evk_gen(sk, dbc, 1, 77)
    evk.tofile(fkey)
    r = subprocess.run([exe, str(n), fin, fout, fkey, str(dbc), str(ic), str(fc)], capture_output=True, text=True, timeout=1800)
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
    # Evaluator::set_relinearization: the same graphs with every product relinearised
    s_rl = sin_relin_oracle(o, ct(0), ct(8), evk, dbc)
    step_rl = approximated_step_relin_oracle(o, ct(5), ct(6), ct(7), 8, 2, 0.25, 2, data[8 * W:16 * W].copy(), evk, dbc)
    want += [s_rl, step_rl]
    want = np.concatenate(want)
    assert got.size == want.size
    assert (got == want).all()
    # (1 - t) A + t B = 0.375 * 12 + 0.625 * 200: the size-3 result of the fused Linear decrypts
    assert abs(o.decode(o.decrypt(sk, lin, 3)) - 129.5) < 1e-6
    if n == 8192:
        # four limbs carry both multiplicative levels of Cubic: the size-4 level-1 result decrypts to
        # ((a + b) t^2 + c t) / 2 + B of the plaintext mirror (t3 = t * t, fhe_resize.h:175)
        A, B, C, D, t = 12.0, 200.0, 131.0, 77.0, 0.625
        a, b, cc = -A + 3 * B - 3 * C + D, 2 * A - 5 * B + 4 * C - D, C - A
        assert abs(o.decode(o.decrypt(sk, c1, 4)) - ((a * t * t + b * t * t + cc * t) * 0.5 + B)) < 1e-6
    assert abs(o.decode(o.decrypt(sk, tz, 2)) - 0.375) < 1e-9
