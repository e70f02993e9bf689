"""The reference's own known answer at the Evaluator boundary, restated.

/root/reference/tests/parameters.cpp:100-128 (SEAL's parameter-selection example, kept by the
reference as its only arithmetic self-check) encrypts x = 12345 with IntegerEncoder(t, base 3),
evaluates 42 x^3 - 27 x + 1 homomorphically -- exponentiate(x, 3, evk) = multiply + relinearize
twice (dbc 15, :89-90), multiply_plain(42), multiply_plain(27), sub, add_plain(1) -- decrypts and
prints 79017370138936.  The value is reproduced here (a) by the CPU oracle alone and (b) by the CUDA
library through the C-ABI on oracle-made encryptions, with (b) also bit-identical to (a).  The
parameters SEAL's ChooserEvaluator would pick are not derivable without SEAL; n = 4096 with
coeff_modulus_128(4096) and t = 1 << 14 (the reference's application parameters, fhe_image.h:22-28)
hold the plaintext growth (|coefficients| < 500 << t/2) and the noise.
"""
import numpy as np
import pytest

X, ANSWER, DBC = 12345, 79017370138936, 15
assert 42 * X ** 3 - 27 * X + 1 == ANSWER


def int_encode(value: int, n: int, t: int, base: int = 3) -> np.ndarray:
    """SEAL IntegerEncoder, odd base: balanced digits in (-base/2, base/2], negative digits as t - |d|."""
    out = np.zeros(n, dtype=np.uint64)
    sign, v, i = (1 if value >= 0 else -1), abs(value), 0
    while v:
        d = v % base
        v //= base
        if d > base // 2:
            d -= base
            v += 1
        d *= sign
        out[i] = d % t
        i += 1
    return out


def int_decode(plain, t: int, base: int = 3) -> int:
    acc = 0
    for c in reversed([int(c) for c in plain]):
        acc = acc * base + (c if c <= t // 2 else c - t)
    return acc


def test_integer_encoder_restatement():
    t = 1 << 14
    for v in (0, 1, 2, 3, 4, 42, 27, 12345, -12345, 79017370138936):
        p = int_encode(v, 64, t)
        assert int_decode(p, t) == v
        assert all(int(c) in (0, 1, t - 1) for c in p)


def oracle_graph(o, x_ct, evk, p42, p27, p1):
    """parameters.cpp:113-120 on the oracle; returns the size-2 result ciphertext."""
    x2 = o.relinearize(o.multiply(x_ct, 2, x_ct, 2), 3, evk, DBC, 1)
    x3 = o.relinearize(o.multiply(x2, 2, x_ct, 2), 3, evk, DBC, 1)
    o.multiply_plain(x3, p42)
    d1 = x_ct.copy()
    o.multiply_plain(d1, p27)
    o.sub(x3, d1)
    o.add_plain(x3, p1)
    return x3


@pytest.fixture(scope="module")
def setup(oracle_mod):
    n, t = 4096, 1 << 14
    o = oracle_mod.Oracle(n, oracle_mod.default_coeff_modulus(n), t, 64, 32)
    sk, pk = o.keygen(2018)
    evk = o.evk_gen(sk, DBC, 1, 11)
    x_ct = o.encrypt(pk, int_encode(X, n, t), 5)
    plains = [int_encode(v, n, t) for v in (42, 27, 1)]
    return o, sk, evk, x_ct, plains


def test_polynomial_known_answer_oracle(setup):
    o, sk, evk, x_ct, (p42, p27, p1) = setup
    res = oracle_graph(o, x_ct.copy(), evk, p42, p27, p1)
    plain, budget = o.decrypt(sk, res, 2, True)
    assert int_decode(plain, o.t) == ANSWER
    assert budget > 10


@pytest.mark.gpu
def test_polynomial_known_answer_cuda(setup, pkg):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    pkg.build()
    o, sk, evk, x_ct, (p42, p27, p1) = setup
    ctx = pkg.Context(o.n, o.q, o.t)
    keys = pkg.EvaluationKeys(ctx, evk, DBC, 1)
    tab = pkg.PlainTable(ctx, np.concatenate([p42, p27, p1]).reshape(3, o.n))
    dx = ctx.to_device(x_ct)
    d3, dx2, dx3 = ctx.alloc(ctx.ct_bytes(3)), ctx.alloc(ctx.ct_bytes(2)), ctx.alloc(ctx.ct_bytes(2))
    ctx.multiply(dx, 2, dx, 2, d3)
    ctx.relinearize(d3, 3, keys, dx2)
    ctx.multiply(dx2, 2, dx, 2, d3)
    ctx.relinearize(d3, 3, keys, dx3)
    ctx.multiply_plain(dx3, 1, tab, 0)
    d1 = ctx.to_device(x_ct)
    ctx.multiply_plain(d1, 1, tab, 1)
    ctx.sub(dx3, d1, 1)
    ctx.add_plain(dx3, 1, tab, 2)
    got = np.empty(ctx.ct_words(2), dtype=np.uint64)
    ctx.download(got, dx3)
    want = oracle_graph(o, x_ct.copy(), evk, p42, p27, p1)
    assert (got == want).all(), "CUDA evaluation of 42x^3-27x+1 differs from the oracle"
    assert int_decode(o.decrypt(sk, got, 2), o.t) == ANSWER
    tab.close()
    keys.close()
    for d in (dx, d3, dx2, dx3, d1):
        ctx.free(d)
    ctx.close()
