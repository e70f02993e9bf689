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


def sincos_relin_oracle(o, x, zero, evk, dbc, sign):
    """homomorphic_sin (sign +1, fhe_decode.h:48-120) / homomorphic_cos (sign -1, :128-200) op by op on the
    oracle with every ct x ct product relinearised to size 2 (host/seal_b200.hpp
    Evaluator::set_relinearization)."""
    import math

    def rl(c):
        return o.relinearize(c, 3, evk, dbc, 1)

    def sq(c):
        return rl(o.square(c, 2))

    def mul(a, b):
        return rl(o.multiply(a, 2, b, 2))

    sh = x.copy()
    o.add_plain(sh, o.encode(-3 * math.pi / 2.0))
    p2 = sq(sh)
    p4 = sq(sq(sh))
    p6 = mul(mul(sq(sq(sh)), sh), sh)
    p8 = sq(sq(sq(sh)))
    p10 = mul(mul(sq(sq(sq(sh))), sh), sh)
    res = zero.copy()
    o.add_plain(res, o.encode(sign * -1.0))
    for term, coef in ((p2, 0.5), (p4, -1 / 24.0), (p6, 1 / 720.0), (p8, -1 / 40320.0), (p10, 1 / 3628800.0)):
        o.multiply_plain(term, o.encode(sign * coef))
        o.add(res, term)
    return res


def sin_relin_oracle(o, x, zero, evk, dbc):
    return sincos_relin_oracle(o, x, zero, evk, dbc, 1.0)


def approximated_step_relin_oracle(o, amplitude, index, count, order, degree, delta, npos, zeros, evk, dbc):
    """approximated_step (fhe_decode.h:202-242) op by op -- `offset += encode(i)` inside the j loop and
    the float casts included -- with every product relinearised; `zeros`: the fresh encryptions the
    series start from, consumed sin then cos per (i, j)."""
    import math
    W = o.ct_words(2)
    b = count.copy()
    o.multiply_plain(b, o.encode(0.5))
    offset = index.copy()
    o.add(offset, b)
    o.add_plain(offset, o.encode(-0.5))
    o.negate(offset)
    o.add_plain(b, o.encode(delta - 0.5))
    run, nz = [], 0
    for i in range(npos):
        c = b.copy()
        o.multiply_plain(c, o.encode(1.0 / float(order)))
        for j in range(1, degree + 1):
            arg_factor = float(np.float32(j)) * math.pi / float(order)
            sin_arg = b.copy()
            o.multiply_plain(sin_arg, o.encode(arg_factor))
            cos_arg = offset.copy()
            o.add_plain(offset, o.encode(float(i)))
            o.multiply_plain(cos_arg, o.encode(arg_factor))
            s = sincos_relin_oracle(o, sin_arg, zeros[nz * W:(nz + 1) * W], evk, dbc, 1.0)
            cs = sincos_relin_oracle(o, cos_arg, zeros[(nz + 1) * W:(nz + 2) * W], evk, dbc, -1.0)
            nz += 2
            term = o.relinearize(o.multiply(s, 2, cs, 2), 3, evk, dbc, 1)
            o.multiply_plain(term, o.encode(2.0 / (math.pi * float(np.float32(j)))))
            o.add(c, term)
        c = o.relinearize(o.multiply(c, 2, amplitude.copy(), 2), 3, evk, dbc, 1)
        run.append(c)
    return np.concatenate(run)


def cubic_relin_oracle(o, A, B, C, D, t, evk, dbc):
    """Cubic (fhe_resize.h:143-189, incl. t3 = t * t) with every product relinearised."""
    def rl(c):
        return o.relinearize(c, 3, evk, dbc, 1)

    def mp(x, v):
        y = x.copy()
        o.multiply_plain(y, o.encode(float(v)))
        return y

    a = mp(B, 3)
    o.sub(a, A)
    o.sub(a, mp(C, 3))
    o.add(a, D)
    b = mp(A, 2)
    o.sub(b, mp(B, 5))
    o.add(b, mp(C, 4))
    o.sub(b, D)
    c = C.copy()
    o.sub(c, A)
    t2 = rl(o.square(t.copy(), 2))
    t3 = rl(o.multiply(t.copy(), 2, t.copy(), 2))
    a = rl(o.multiply(a, 2, t3, 2))
    b = rl(o.multiply(b, 2, t2, 2))
    c = rl(o.multiply(c, 2, t.copy(), 2))
    o.add(a, b)
    o.add(a, c)
    o.multiply_plain(a, o.encode(0.5))
    o.add(a, B)
    return a


def linear_relin_oracle(o, A, B, t, evk, dbc):
    """Linear (fhe_resize.h:191-204): (1 - t) A + t B via negate, add_plain(1.0), two multiplies, add."""
    one_minus_t = t.copy()
    o.negate(one_minus_t)
    o.add_plain(one_minus_t, o.encode(1.0))
    a = o.relinearize(o.multiply(one_minus_t, 2, A.copy(), 2), 3, evk, dbc, 1)
    b = o.relinearize(o.multiply(B.copy(), 2, t.copy(), 2), 3, evk, dbc, 1)
    o.add(a, b)
    return a


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
