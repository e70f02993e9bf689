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


# SEAL 2.3.1 internal_mods ([SEAL-recall]): largest primes below 2^61 that are
# 1 mod 2^18, descending: m_sk, gamma, then the auxiliary base.
M_SK = 0x1FFFFFFFFFE00001
AUX_SMALL_MODS = [0x1FFFFFFFFFB40001, 0x1FFFFFFFFF500001, 0x1FFFFFFFFF380001, 0x1FFFFFFFFF000001,
                  0x1FFFFFFFFEF00001, 0x1FFFFFFFFEE80001, 0x1FFFFFFFFEB40001, 0x1FFFFFFFFE780001,
                  0x1FFFFFFFFE600001, 0x1FFFFFFFFE4C0001]
M_TILDE = 1 << 32


class Behz:
    """BEHZ full-RNS multiply on Python ints: every fast base conversion
    is the exact big-integer sum  S = sum_i y_i (q/q_i)  reduced afterwards, the Montgomery step is
    an exact integer division -- no Barrett, no NTT, no 128-bit accumulators, so nothing is shared
    with oracle/fhe_oracle_behz.c or the CUDA kernels beyond the specification itself."""

    def __init__(self, m: Model, bsk: list[int] | None = None):
        """bsk: another auxiliary base B + [m_sk] (the CUDA library uses 55-bit primes, csrc/behz.cuh);
        default: SEAL's.  The products do not depend on it (tests/test_oracle_behz.py)."""
        self.m = m
        k = len(m.qs)
        bits = sum(q.bit_length() for q in m.qs)
        kb = k + (1 if 32 + m.t.bit_length() + bits >= 61 * k + 61 else 0)
        self.B = list(bsk[:-1]) if bsk else AUX_SMALL_MODS[:kb]
        self.msk = bsk[-1] if bsk else M_SK
        self.bsk = self.B + [self.msk]
        self.Bprod = prod(self.B)

    def _conv_sum(self, res, mods, total):
        """sum_i [res_i (M/m_i)^-1]_{m_i} (M/m_i): the uncorrected fast base conversion."""
        s = 0
        for r, mi in zip(res, mods):
            hat = total // mi
            s += (r * pow(hat % mi, -1, mi) % mi) * hat
        return s

    def extend(self, poly):
        """poly: n ints mod q -> [len(bsk)][n] residues (fastbconv_mtilde + mont_rq)."""
        m, q = self.m, self.m.q
        out = [[0] * m.n for _ in self.bsk]
        for c, x in enumerate(poly):
            s = self._conv_sum([(x % qi) * M_TILDE % qi for qi in m.qs], m.qs, q)
            r = (-s * pow(q, -1, M_TILDE)) % M_TILDE
            if r >= M_TILDE // 2:
                r -= M_TILDE
            v, rem = divmod(s + q * r, M_TILDE)
            assert rem == 0
            for j, p in enumerate(self.bsk):
                out[j][c] = v % p
        return out

    def scale_round(self, xq, xb):
        """xq: [k][n] residues, xb: [len(bsk)][n] -> n ints mod q (times t, fast_floor, fastbconv_sk)."""
        m, q, t = self.m, self.m.q, self.m.t
        out = []
        for c in range(m.n):
            s = self._conv_sum([t * xq[i][c] % qi for i, qi in enumerate(m.qs)], m.qs, q)
            f = [(t * xb[j][c] - s) * pow(q, -1, p) % p for j, p in enumerate(self.bsk)]
            s2 = self._conv_sum(f[:-1], self.B, self.Bprod)
            alpha = (s2 - f[-1]) * pow(self.Bprod, -1, self.msk) % self.msk
            if alpha > self.msk // 2:
                alpha -= self.msk
            out.append((s2 - alpha * self.Bprod) % q)   # CRT of the per-limb results
        return out

    def multiply(self, a, b):
        """a, b: lists of polys (ints mod q) -> list of len(a)+len(b)-1 polys."""
        m = self.m
        ea, eb = [self.extend(p) for p in a], [self.extend(p) for p in b]
        out = []
        for l in range(len(a) + len(b) - 1):
            accq = [[0] * m.n for _ in m.qs]
            accb = [[0] * m.n for _ in self.bsk]
            for x in range(len(a)):
                y = l - x
                if y < 0 or y >= len(b):
                    continue
                for i, qi in enumerate(m.qs):
                    pr = m.poly_mul([v % qi for v in a[x]], [v % qi for v in b[y]], qi)
                    accq[i] = [(u + v) % qi for u, v in zip(accq[i], pr)]
                for j, p in enumerate(self.bsk):
                    pr = m.poly_mul(ea[x][j], eb[y][j], p)
                    accb[j] = [(u + v) % p for u, v in zip(accb[j], pr)]
            out.append(self.scale_round(accq, accb))
        return out

    def relinearize(self, ct, evk, dbc):
        """evk[r][i][l] = (k0, k1) polys mod q; SURVEY A.7."""
        m = self.m
        ct = [list(p) for p in ct]
        while len(ct) > 2:
            last, r = ct.pop(), len(ct) - 2
            for i, qi in enumerate(m.qs):
                for l, (k0, k1) in enumerate(evk[r][i]):
                    d = [((c % qi) >> (l * dbc)) & ((1 << dbc) - 1) for c in last]
                    ct[0] = m.poly_add(ct[0], m.poly_mul(d, k0))
                    ct[1] = m.poly_add(ct[1], m.poly_mul(d, k1))
        return ct
