"""Compositions of the CPU oracle's primitives into the reference's op graphs where liboracle.so has
no single entry point for them: the bicubic ResizeImage loop over orc_cubic, and the graphs with every
ciphertext product relinearised (Cubic, Linear, the Taylor series, the step approximation).

TEST INFRASTRUCTURE ONLY -- see oracle/fhe_oracle.h.  Used by tests/ and by bench.py's cpu_baseline /
parity legs; the product package never imports this module.  `o` is an oracle_py.Oracle."""
import numpy as np


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
