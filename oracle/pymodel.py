"""Independent pure-Python big-int model of the BFV ciphertext arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/fhe_oracle.h).  PARITY UNPINNED: SEAL 2.3 is
absent from /root/reference; this model restates the published scheme in the
simplest possible form -- one big modulus q = prod q_j, schoolbook negacyclic
products on Python ints, exact rational rounding -- so that it shares no code or
technique (no NTT, no RNS, no Barrett) with oracle/fhe_oracle.c or the CUDA path.
Agreement of the two is the substitute for running SEAL itself.

Ring: R_q = Z_q[x]/(x^n+1).  A ciphertext is a list of `size` polynomials, each a
list of n Python ints in [0, q).  to_rns()/from_rns() convert to the
[poly][limb][n] layout used by the oracle and the device.

Also holds the harness-side scheme pieces the reference keeps on the client
(client_jpeg.cpp:96-116,140-165,266-279): keygen, encrypt, decrypt.
"""
from __future__ import annotations

import random
from fractions import Fraction
from math import prod


class Model:
    def __init__(self, n: int, q: list[int], t: int):
        self.n, self.qs, self.t = n, list(q), t
        self.q = prod(q)
        self.delta = self.q // t           # SEAL "coeff_div_plain_modulus"
        self.q_mod_t = self.q % t          # SEAL "upper_half_increment"
        self.thr = (t + 1) // 2            # plain_upper_half_threshold

    # ---------------------------------------------------------------- ring
    def poly_mul(self, a, b, mod=None):
        """Schoolbook negacyclic product."""
        n, m = self.n, (mod or self.q)
        out = [0] * n
        for i, ai in enumerate(a):
            if not ai:
                continue
            for j, bj in enumerate(b):
                if not bj:
                    continue
                d = i + j
                if d >= n:
                    out[d - n] -= ai * bj
                else:
                    out[d] += ai * bj
        return [v % m for v in out]

    def poly_add(self, a, b): return [(x + y) % self.q for x, y in zip(a, b)]
    def poly_sub(self, a, b): return [(x - y) % self.q for x, y in zip(a, b)]
    def poly_neg(self, a): return [(-x) % self.q for x in a]

    # ---------------------------------------------------------------- RNS layout
    def to_rns(self, ct):
        """[poly][n] big ints -> flat list in [poly][limb][n] order."""
        return [c % qj for poly in ct for qj in self.qs for c in poly]

    def from_rns(self, flat, size):
        n, k = self.n, len(self.qs)
        # CRT recombination
        Mi = [self.q // qj for qj in self.qs]
        yi = [pow(Mi[j] % self.qs[j], -1, self.qs[j]) for j in range(k)]
        ct = []
        for p in range(size):
            poly = []
            for i in range(n):
                v = 0
                for j in range(k):
                    r = int(flat[(p * k + j) * n + i])
                    v += r * yi[j] % self.qs[j] * Mi[j]
                poly.append(v % self.q)
            ct.append(poly)
        return ct

    # ---------------------------------------------------------------- plaintext handling
    def lift(self, plain):
        """Centred lift of m in Z_t into Z_q."""
        out = [0] * self.n
        for i, m in enumerate(plain):
            m = int(m)
            out[i] = m if m < self.thr else (m - self.t) % self.q
        return out

    def scale(self, plain):
        """Delta-scaling used by encrypt/add_plain/sub_plain."""
        out = [0] * self.n
        for i, m in enumerate(plain):
            m = int(m)
            out[i] = (self.delta * m + (self.q_mod_t if m >= self.thr else 0)) % self.q
        return out

    # ---------------------------------------------------------------- Evaluator subset
    def add(self, a, b):
        size = max(len(a), len(b))
        z = [0] * self.n
        return [self.poly_add(a[p] if p < len(a) else z, b[p] if p < len(b) else z) for p in range(size)]

    def sub(self, a, b):
        size = max(len(a), len(b))
        z = [0] * self.n
        return [self.poly_sub(a[p] if p < len(a) else z, b[p] if p < len(b) else z) for p in range(size)]

    def negate(self, a): return [self.poly_neg(p) for p in a]
    def add_plain(self, a, plain): return [self.poly_add(a[0], self.scale(plain))] + [list(p) for p in a[1:]]
    def sub_plain(self, a, plain): return [self.poly_sub(a[0], self.scale(plain))] + [list(p) for p in a[1:]]

    def multiply_plain(self, a, plain):
        if not any(int(m) for m in plain):
            raise ValueError("plain cannot be zero")
        lp = self.lift(plain)
        return [self.poly_mul(p, lp) for p in a]

    # ---------------------------------------------------------------- FractionalEncoder
    def encode(self, value: float, frac_coeffs: int = 100):
        """FractionalEncoder(t, n, int, frac, base 2).encode, computed with
        exact rationals instead of repeated doubling of a double."""
        n, t = self.n, self.t
        v = Fraction(value)          # exact value of the double
        sign = -1 if v < 0 else 1
        mag = abs(v)
        ipart = mag.numerator // mag.denominator
        frac = mag - ipart
        out = [0] * n
        i = 0
        while ipart:
            if ipart & 1:
                out[i] = sign % t
            ipart >>= 1
            i += 1
        for d in range(frac_coeffs):
            frac *= 2
            if frac >= 1:
                frac -= 1
                out[n - 1 - d] = (-sign) % t
        return out

    def decode(self, plain) -> float:
        """Centred coefficients; x^i -> 2^i (i < n/2), x^(n-i) -> -2^-i."""
        n, acc = self.n, Fraction(0)
        for i, m in enumerate(plain):
            m = int(m)
            if not m:
                continue
            c = m if m < self.thr else m - self.t
            acc += c * Fraction(2) ** i if i < n // 2 else -c * Fraction(1, 2 ** (n - i))
        return float(acc)

    # ---------------------------------------------------------------- harness: keygen/encrypt/decrypt
    def keygen(self, rng: random.Random, mul=None):
        """sk ternary; pk = (-(a s + e), a)."""
        mul = mul or self.poly_mul
        n, q = self.n, self.q
        s = [rng.choice((-1, 0, 1)) % q for _ in range(n)]
        a = [rng.randrange(q) for _ in range(n)]
        e = self._noise(rng)
        pk0 = self.poly_neg(self.poly_add(mul(a, s), e))
        return s, (pk0, a)

    def _noise(self, rng, sigma=3.19):
        out = []
        for _ in range(self.n):
            while True:
                x = round(rng.gauss(0, sigma))
                if abs(x) <= 6 * sigma:
                    break
            out.append(x % self.q)
        return out

    def encrypt(self, plain, pk, rng: random.Random, mul=None):
        mul = mul or self.poly_mul
        u = [rng.choice((-1, 0, 1)) % self.q for _ in range(self.n)]
        c0 = self.poly_add(self.poly_add(mul(pk[0], u), self._noise(rng)), self.scale(plain))
        c1 = self.poly_add(mul(pk[1], u), self._noise(rng))
        return [c0, c1]

    def decrypt(self, ct, s, mul=None, return_noise=False):
        """[c0 + c1 s + c2 s^2 ...]_q -> round(t x / q) mod t, exact rational rounding."""
        mul = mul or self.poly_mul
        acc, sp = list(ct[0]), None
        for p in ct[1:]:
            sp = s if sp is None else mul(sp, s)
            acc = self.poly_add(acc, mul(p, sp))
        q, t = self.q, self.t
        out, worst = [], Fraction(0)
        for x in acc:
            num = t * x
            m = (2 * num + q) // (2 * q)   # round half up
            if return_noise:
                err = abs(Fraction(num, q) - m)
                worst = max(worst, err)
            out.append(m % t)
        return (out, worst) if return_noise else out
