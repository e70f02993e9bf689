// modarith.cuh -- 64-bit modular arithmetic over the RNS limbs (q_j < 2^55).
//
// B200 replacement for SEAL 2.3's util/uintarithsmallmod (Barrett) and the Harvey/Shoup lazy
// multiplication inside util/smallntt (both [SEAL-recall], sources absent).
//
// Cost model measured on B200 (profiles/r01_*): the FMA pipe issues a 32-bit IMAD per cycle per
// SM sub-partition and an IMAD.WIDE / IMAD.HI every 2 cycles; the ALU pipe (IADD3, SEL, SHF,
// ISETP) issues every 2 cycles.  A textbook Shoup multiplication (exact mul.hi.u64 + two
// mul.lo.u64) is 17 SASS instructions.  `mulc` below is 11: the quotient estimate drops the
// low x low partial product and keeps the carries out of the critical path by scaling the
// companion constant with 2^63 instead of 2^64, at the price of a wider lazy range [0, 5q).
// All lazy ranges are tracked at compile time by the callers (ntt.cuh, dct.cuh); every value
// handed back to the user is canonical ([0, q)), so results are bit-identical to SEAL's.
//
// Every function is __host__ __device__ with one arithmetic definition (the device path only
// spells the same operations as PTX), so tests/emu executes the exact kernel bodies on the CPU.
#pragma once
#include <cstdint>

typedef unsigned long long u64;
typedef unsigned int u32;

#if defined(__CUDACC__)
#define FHE_HD __host__ __device__ __forceinline__
#else
#define FHE_HD inline
#endif

// A multiplier constant w in [0, q) with its companion p = floor(w * 2^63 / q) < 2^63.
struct alignas(16) W2 {
    u64 w, p;
};

// Per-limb constants of the lazy arithmetic.
struct LimbArith {
    u64 q;     // the prime, < 2^55
    u64 n2q;   // -2q mod 2^64
    u64 nq;    // -q mod 2^64
    u64 q5;    // 5q: offset that keeps  x + q5 - mulc(..)  non-negative
    u64 kq5[4];   // 5q * 2^i  and
    u64 kq16[4];  // 16q * 2^i: the offsets of the lazy subtractions (ntt.cuh inverse stages), as
                  // ready-made 64-bit operands -- computing bound * q in the kernel costs an
                  // IMAD.WIDE on the FMA pipe per butterfly (measured: inverse NTT 14 % slower)
    u32 mu63;  // floor(2^63 / q): quotient estimate of reduce63
    u32 mu64;  // floor(2^64 / q): quotient estimate of canon
};

// b * q for the bounds the lazy schedules produce (5 * 2^i, 16 * 2^i, i < 4); b is a compile-time
// constant at every call site, so the chain folds to one operand.
FHE_HD u64 bound_times_q(const LimbArith &L, int b)
{
    return b == 5 ? L.kq5[0] : b == 10 ? L.kq5[1] : b == 20 ? L.kq5[2] : b == 40 ? L.kq5[3]
         : b == 16 ? L.kq16[0] : b == 32 ? L.kq16[1] : b == 64 ? L.kq16[2] : b == 128 ? L.kq16[3]
         : (u64)b * L.q;
}

FHE_HD u64 mulhi_u64(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
FHE_HD u32 mulhi_u32(u32 a, u32 b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}

// x * c.w - Q * 2q  with  Q = x1*p1 + floor((x0*p1 + x1*p0) / 2^32)  (x = x1:x0, p = p1:p0).
// Requires x < 2^63, c.w < q < 2^55.  Q is within 2.5 below x*w/(2q) and never above it, so the
// result is congruent to x*w mod q and lies in [0, 5q).  x0*p1 + x1*p0 < 2^64 because both p
// and x are below 2^63: there is no carry to propagate.
FHE_HD u64 mulc(u64 x, W2 c, u64 n2q)
{
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32), w0 = (u32)c.w, w1 = (u32)(c.w >> 32);
    const u32 p0 = (u32)c.p, p1 = (u32)(c.p >> 32), n0 = (u32)n2q, n1 = (u32)(n2q >> 32);
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u32 h, z, q0, q1, lo, hi;\n\t"
        ".reg .u64 t, H, Q, acc;\n\t"
        "mul.wide.u32 t, %2, %5;\n\t"          // x1*p0
        "mad.wide.u32 t, %1, %6, t;\n\t"       // + x0*p1
        "mov.b64 {z, h}, t;\n\t"
        "cvt.u64.u32 H, h;\n\t"
        "mad.wide.u32 Q, %2, %6, H;\n\t"       // Q = x1*p1 + h
        "mov.b64 {q0, q1}, Q;\n\t"
        "mul.wide.u32 acc, %1, %3;\n\t"        // x0*w0
        "mad.wide.u32 acc, q0, %7, acc;\n\t"   // + Q0*n0
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, %1, %4, hi;\n\t"       // + 2^32 (x0*w1 + x1*w0 + Q0*n1 + Q1*n0)
        "mad.lo.u32 hi, %2, %3, hi;\n\t"
        "mad.lo.u32 hi, q0, %8, hi;\n\t"
        "mad.lo.u32 hi, q1, %7, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "=l"(r)
        : "r"(x0), "r"(x1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0), "r"(n1));
    return r;
#else
    const u64 t = (u64)x0 * p1 + (u64)x1 * p0;
    const u64 Q = (u64)x1 * p1 + (t >> 32);
    return x * c.w + Q * n2q;
#endif
}

// Sums of lazy products, reduced once.  mulc(x_i, c_i) = x_i w_i - Q_i 2q mod 2^64 and a sum of m of them
// is below 5 m q < 2^64, so  sum_i x_i w_i - (sum_i Q_i) 2q mod 2^64  IS that sum, bit for bit -- with ONE
// product Q * 2q (1 wide + 2 narrow multiplies) for the whole sum instead of one per term: a term costs
// 4 wide + 2 narrow multiplies instead of 5 + 4.  Used where a value is a sum of products and nothing
// else reads the individual products (relinearisation sums, the DCT butterflies' odd part, the colour
// matrix).  MulcSum is the pair (sum of low words, sum of quotient estimates), both mod 2^64.
struct MulcSum {
    u64 lo, Q;
};
FHE_HD MulcSum mulc_part(u64 x, W2 c)
{
    const u32 x0 = (u32)x, x1 = (u32)(x >> 32), p0 = (u32)c.p, p1 = (u32)(c.p >> 32);
    const u64 t = (u64)x1 * p0 + (u64)x0 * p1;
    MulcSum s;
    s.Q = (u64)x1 * p1 + (t >> 32);
    s.lo = x * c.w;
    return s;
}
FHE_HD MulcSum operator+(MulcSum a, MulcSum b)
{
    MulcSum s;
    s.lo = a.lo + b.lo;
    s.Q = a.Q + b.Q;
    return s;
}
FHE_HD u64 mulc_done(MulcSum s, u64 n2q) { return s.lo + s.Q * n2q; }

// x < 2^63  ->  congruent value in [0, 4q):  x - Q*2q with Q = floor(x1 * mu63 / 2^32), which
// is floor(x/(2q)) or one less (mu63 = floor(2^63/q) < 2^32 because q > 2^31).
FHE_HD u64 reduce63(u64 x, const LimbArith &L)
{
    const u32 Q = mulhi_u32((u32)(x >> 32), L.mu63);
    return x + (u64)Q * L.n2q;  // Q < 2^32: IMAD.WIDE + IMAD
}

// x in [0, 2m) -> [0, m), for m < 2^63 (the sign bit of x - m says which)
FHE_HD u64 csub(u64 x, u64 m)
{
    const u64 t = x - m;
    return (long long)t < 0 ? x : t;
}

// x < 2^63  ->  x mod q, canonical.  k = floor(x1 * mu64 / 2^32) is floor(x/q) or one less.
FHE_HD u64 canon(u64 x, const LimbArith &L)
{
    const u32 k = mulhi_u32((u32)(x >> 32), L.mu64);
    return csub(x + (u64)k * L.nq, L.q);
}

FHE_HD u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
FHE_HD u64 submod(u64 a, u64 b, u64 q) { return csub(a - b + q, q); }
FHE_HD u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

// Barrett 128 -> 64 with ratio = floor(2^128 / q) = (r1:r0); result in [0, q).  Table setup only.
FHE_HD u64 barrett128(u64 z0, u64 z1, u64 q, u64 r0, u64 r1)
{
    u64 t0h = mulhi_u64(z0, r0);
    u64 m_lo = z0 * r1, m_hi = mulhi_u64(z0, r1);
    u64 s = m_lo + t0h;
    u64 carry1 = s < m_lo;
    u64 n_lo = z1 * r0, n_hi = mulhi_u64(z1, r0);
    u64 s2 = n_lo + s;
    u64 carry2 = s2 < n_lo;
    u64 qhat = z1 * r1 + m_hi + carry1 + n_hi + carry2;
    u64 res = z0 - qhat * q;
    return csub(res, q);
}
FHE_HD u64 mulmod_barrett(u64 a, u64 b, u64 q, u64 r0, u64 r1)
{
    return barrett128(a * b, mulhi_u64(a, b), q, r0, r1);
}

// SplitMix64 finaliser: the synthetic-input generator shared with the oracle.
FHE_HD u64 splitmix64(u64 z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
