// modarith.cuh -- 64-bit modular arithmetic over the RNS limbs (q_j < 2^62).
//
// B200 replacement for SEAL 2.3's util/uintarithsmallmod (Barrett) and the Harvey/Shoup lazy
// multiplication inside util/smallntt (both [SEAL-recall], sources absent).
// Every function is __host__ __device__ so the exact kernel bodies can also be executed
// thread-by-thread on the CPU by tests/emu (index math is checked here, where there is no GPU).
#pragma once
#include <cstdint>

typedef unsigned long long u64;

#if defined(__CUDACC__)
#define FHE_HD __host__ __device__ __forceinline__
#else
#define FHE_HD inline
#endif

struct W2 {  // twiddle / constant with its Shoup companion floor(w * 2^64 / q)
    u64 w, wp;
};

FHE_HD u64 mulhi_u64(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

FHE_HD u64 umin_u64(u64 a, u64 b) { return a < b ? a : b; }

// x * w mod q, result in [0, 2q), valid for ANY 64-bit x (w < q, wp = floor(w 2^64 / q)).
FHE_HD u64 shoup_lazy(u64 x, u64 w, u64 wp, u64 q) { return x * w - mulhi_u64(x, wp) * q; }
FHE_HD u64 shoup_lazy(u64 x, W2 c, u64 q) { return shoup_lazy(x, c.w, c.wp, q); }

// x in [0, 2m) -> [0, m)   (x - m wraps to a huge value when x < m, min picks x)
FHE_HD u64 csub(u64 x, u64 m) { return umin_u64(x, x - m); }

FHE_HD u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
FHE_HD u64 submod(u64 a, u64 b, u64 q) { return csub(a - b + q, q); }
FHE_HD u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

// Barrett 128 -> 64 with ratio = floor(2^128 / q) = (r1:r0); result in [0, q).
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
