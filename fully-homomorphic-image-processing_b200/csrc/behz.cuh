// behz.cuh -- Evaluator::multiply / square (BFV tensor product with the BEHZ full-RNS scaling) and
// Evaluator::relinearize on the device.
//
// Replaces SEAL 2.3's Evaluator::multiply / square / relinearize and util/baseconverter
// (fastbconv_mtilde, mont_rq, fast_floor, fastbconv_sk) -- [SEAL-recall], sources absent; the
// procedure is SURVEY.md Appendix A.6 / A.7 and oracle/fhe_oracle_behz.c restates it step by step.
// Call sites in the reference: fhe_resize.h:174-179,197-198 (Cubic, Linear), fhe_decode.h:67-97,
// 147-177,235,239 (sin / cos / step).
//
// Every step is exact arithmetic modulo a prime with canonical results, so constants may be merged
// freely: here m~^-1 is folded into the base-conversion matrix of the extension, t and q^-1 into
// the matrices of fast_floor, B^-1 into the m_sk row of the Shenoy-Kumaresan correction.  Each of
// the two per-coefficient kernels then is: k (or |B|) multiplications by a constant, and one
// 128-bit multiply-accumulate row with a single Barrett reduction per output residue.
//
// The auxiliary base.  SEAL uses its 61-bit internal primes for B_sk.  The RESULT of the procedure does
// not depend on that choice: the extension yields the residues of one integer X' = (S + q r) / m~
// fixed by the q-residues alone, fast_floor the residues of one integer W = (t P - S') / q, and the
// Shenoy-Kumaresan step returns W mod q exactly as long as |W| < B m_sk / 2 -- any base that is large
// enough gives the same bits (tests/test_oracle_behz.py checks the big-int model with both bases;
// every GPU parity test compares this base against the oracle's SEAL base).  So this library takes
// its auxiliary primes just below 2^55, where the lazy arithmetic of ntt.cuh applies: the B_sk limbs
// run the same butterflies, kernels and tables as the q limbs (no 61-bit transform at all), and the
// primes have the shape 2^55 - delta (delta < 2^26), which turns every 128 -> 64-bit reduction modulo
// them into three small multiplications (fold_aux).
//
// Data layout: a polynomial in base q is [k][n]; in base B_sk = B u {m_sk} it is [nb][n], m_sk last.
#pragma once
#include "ntt_generic.cuh"

namespace fhe {

constexpr int BEHZ_MAXK = 8;    // limbs of q supported by the multiply path
constexpr int BEHZ_MAXB = 10;   // |B_sk| <= 10

typedef unsigned __int128 u128d;

struct BehzDev {
    uint32_t n, logn, k, kb, nb;
    u64 q[BEHZ_MAXK], qr0[BEHZ_MAXK], qr1[BEHZ_MAXK];       // primes of q and floor(2^128 / q_i)
    u32 qdelta[BEHZ_MAXK];                // 2^55 - q_i when that is in (0, 2^26) (fold_delta_of), else 0: Barrett
    u64 b[BEHZ_MAXB], br0[BEHZ_MAXB], br1[BEHZ_MAXB];       // B_sk
    u32 bdelta[BEHZ_MAXB];                // 2^55 - b_j (< 2^26): the auxiliary primes are 2^55 - delta
    // extension q -> B_sk (fastbconv_mtilde + mont_rq)
    W2 mt_qhat_inv[BEHZ_MAXK];            // m~ (q/q_i)^-1 mod q_i         {w, floor(w 2^64 / q_i)}
    u32 qhat_mod_mt[BEHZ_MAXK];           // (q/q_i) mod 2^32
    u32 neg_inv_q_mod_mt;                 // -q^-1 mod 2^32
    u64 C[BEHZ_MAXB][BEHZ_MAXK];          // (q/q_i) m~^-1 mod b_j
    u64 Dpos[BEHZ_MAXB], Dneg[BEHZ_MAXB]; // +- q m~^-1 mod b_j
    // scaling: times t, fast_floor, fastbconv_sk
    W2 T[BEHZ_MAXK];                      // t (q/q_i)^-1 mod q_i
    u64 E[BEHZ_MAXB];                     // t q^-1 mod b_j
    u64 F[BEHZ_MAXB][BEHZ_MAXK];          // -(q/q_i) q^-1 mod b_j
    W2 bhat_inv[BEHZ_MAXB];               // (B/b_l)^-1 mod b_l
    u64 G[BEHZ_MAXB], H;                  // (B/b_l) B^-1 mod m_sk ; -B^-1 mod m_sk
    u64 I[BEHZ_MAXK][BEHZ_MAXB];          // (B/b_l) mod q_i
    u64 Bq_pos[BEHZ_MAXK], Bq_neg[BEHZ_MAXK];  // +- B mod q_i
};

FHE_HD u64 barrett_u128(u128d z, u64 q, u64 r0, u64 r1) { return barrett128((u64)z, (u64)(z >> 64), q, r0, r1); }
// z mod p, canonical, for an auxiliary prime p = 2^55 - delta (delta < 2^26) and z < 2^119 (every
// accumulator here is a sum of at most 33 products of two values below 2^55 and 2^55: < 2^116).
// Three folds of the part above bit 55 (2^55 = delta mod p): 64 x 26, 36 x 26 and 8 x 26-bit products
// instead of the five 64 x 64 products of a Barrett 128 -> 64 step.
FHE_HD u64 fold_aux(u128d z, u64 p, u32 delta)
{
    const u64 M55 = (1ull << 55) - 1;
    const u64 zl = (u64)z & M55, zh = (u64)(z >> 55);                 // zh < 2^64
    const u64 lo = (u64)(u32)zh * delta, hi = (zh >> 32) * delta;       // zh delta = lo + hi 2^32 (< 2^90)
    const u128d w = (u128d)zl + lo + ((u128d)hi << 32);                 // < 2^91
    const u64 wl = (u64)w & M55, wh = (u64)(w >> 55);                   // wh < 2^36
    const u64 s = wl + wh * delta;                                       // < 2^55 + 2^62
    const u64 r = (s & M55) + (s >> 55) * delta;                         // < 2^55 + 2^34 < 2p
    return csub(r, p);
}
// The primes of q themselves usually have that shape: all eight of SEAL's 55-bit defaults are
// 2^55 - a 2^18 + 1 with 2^55 - q < 2^26 (csrc/fhe_b200.cu default chain).  fold_delta_of(q) is that
// delta or 0 ("use Barrett"); reduce_u128 picks the route -- a uniform branch, the prime is per launch or
// per CTA.  Both routes give the canonical residue, so which one runs is unobservable.
FHE_HD u32 fold_delta_of(u64 q)
{
    const u64 top = 1ull << 55;
    return (q < top && top - q < (1ull << 26)) ? (u32)(top - q) : 0u;
}
FHE_HD u64 reduce_u128(u128d z, u64 q, u64 r0, u64 r1, u32 delta)
{
    return delta ? fold_aux(z, q, delta) : barrett_u128(z, q, r0, r1);
}
// x * c mod q, canonical, c a Shoup pair scaled by 2^64
FHE_HD u64 mulsh_canon(u64 x, W2 c, u64 q) { return csub(mulsh(x, c, q), q); }

// ---------------------------------------------------------------------------------------------
// per-coefficient bodies (host + device: tests/emu runs them on the CPU)
// ---------------------------------------------------------------------------------------------
// Both bodies are templates on the base sizes: K = |q| and KB = |B| known at compile time (K > 0:
// every loop unrolls, the residues live in registers) or taken from BehzDev at run time (K == 0).
//
// x: this coefficient's k residues (stride xs); out: nb residues (stride os)
template <int K, int KB>
FHE_HD void behz_extend_coeff_t(const BehzDev &B, const u64 *x, size_t xs, u64 *out, size_t os)
{
    constexpr bool ST = K > 0;
    constexpr int MK = ST ? K : BEHZ_MAXK, MB = ST ? KB + 1 : BEHZ_MAXB;
    const int k = ST ? K : (int)B.k, nb = ST ? KB + 1 : (int)B.nb;
    u64 y[MK];
    u32 vm = 0;
#pragma unroll
    for (int i = 0; i < MK; ++i)
        if (ST || i < k) {
            y[i] = mulsh_canon(x[i * xs], B.mt_qhat_inv[i], B.q[i]);
            vm += (u32)y[i] * B.qhat_mod_mt[i];
        }
    // r = [-v q^-1] mod 2^32, centred in [-2^31, 2^31)
    const u32 r = vm * B.neg_inv_q_mod_mt;
    const bool neg = r >= 0x80000000u;
    const u64 rabs = neg ? (u64)(0u - r) : (u64)r;
#pragma unroll
    for (int j = 0; j < MB; ++j)
        if (ST || j < nb) {
            u128d acc = (u128d)rabs * (neg ? B.Dneg[j] : B.Dpos[j]);
#pragma unroll
            for (int i = 0; i < MK; ++i)
                if (ST || i < k) acc += (u128d)y[i] * B.C[j][i];
            out[j * os] = fold_aux(acc, B.b[j], B.bdelta[j]);
        }
}

// xq: k residues of the tensor result in base q, xb: nb residues in base B_sk; out: k residues
template <int K, int KB>
FHE_HD void behz_scale_coeff_t(const BehzDev &B, const u64 *xq, size_t qs, const u64 *xb, size_t bs, u64 *out, size_t os)
{
    constexpr bool ST = K > 0;
    constexpr int MK = ST ? K : BEHZ_MAXK, MB = ST ? KB + 1 : BEHZ_MAXB;
    const int k = ST ? K : (int)B.k, kb = ST ? KB : (int)B.kb, nb = kb + 1;
    u64 y[MK], f[MB], z[MB];
#pragma unroll
    for (int i = 0; i < MK; ++i)
        if (ST || i < k) y[i] = mulsh_canon(xq[i * qs], B.T[i], B.q[i]);
    // fast_floor: (t x_bsk - conv(t x_q)) q^-1 mod b_j
#pragma unroll
    for (int j = 0; j < MB; ++j)
        if (ST || j < nb) {
            u128d acc = (u128d)xb[j * bs] * B.E[j];
#pragma unroll
            for (int i = 0; i < MK; ++i)
                if (ST || i < k) acc += (u128d)y[i] * B.F[j][i];
            f[j] = fold_aux(acc, B.b[j], B.bdelta[j]);
        }
    // fastbconv_sk: alpha = [(conv_B->msk(f) - f_sk) B^-1] mod m_sk, centred
    u64 fsk = 0;
#pragma unroll
    for (int j = 0; j < MB; ++j)
        if (j == kb) fsk = f[j];
    u128d acc = (u128d)fsk * B.H;
#pragma unroll
    for (int l = 0; l < MB; ++l)
        if (l < kb) {
            z[l] = mulsh_canon(f[l], B.bhat_inv[l], B.b[l]);
            acc += (u128d)z[l] * B.G[l];
        }
    const u64 msk = B.b[kb];
    const u64 alpha = fold_aux(acc, msk, B.bdelta[kb]);
    const bool neg = alpha > (msk >> 1);
    const u64 aabs = neg ? msk - alpha : alpha;
#pragma unroll
    for (int i = 0; i < MK; ++i)
        if (ST || i < k) {
            u128d a2 = (u128d)aabs * (neg ? B.Bq_pos[i] : B.Bq_neg[i]);
#pragma unroll
            for (int l = 0; l < MB; ++l)
                if (l < kb) a2 += (u128d)z[l] * B.I[i][l];
            out[i * os] = reduce_u128(a2, B.q[i], B.qr0[i], B.qr1[i], B.qdelta[i]);
        }
}
// run-time sizes (tests/emu, unusual bases)
FHE_HD void behz_extend_coeff(const BehzDev &B, const u64 *x, size_t xs, u64 *out, size_t os)
{
    behz_extend_coeff_t<0, 0>(B, x, xs, out, os);
}
FHE_HD void behz_scale_coeff(const BehzDev &B, const u64 *xq, size_t qs, const u64 *xb, size_t bs, u64 *out, size_t os)
{
    behz_scale_coeff_t<0, 0>(B, xq, qs, xb, bs, out, os);
}

// one (limb, position) of the tensor product: d[l] = sum_{x+y=l} a[x] b[y] mod q, operands `stride`
// words apart per poly, 128-bit accumulation (min(s1, s2) <= 32 terms of < 2^122), one reduction
// (delta: fold_delta_of(q), or 0 for Barrett)
FHE_HD void behz_tensor_elem(const u64 *a, const u64 *b, u64 *d, size_t stride, uint32_t s1, uint32_t s2,
                             u64 q, u64 r0, u64 r1, u32 delta = 0)
{
    const uint32_t sd = s1 + s2 - 1;
    for (uint32_t l = 0; l < sd; ++l) {
        u128d acc = 0;
        const uint32_t xlo = l >= s2 ? l - (s2 - 1) : 0, xhi = l < s1 ? l : s1 - 1;
        for (uint32_t x = xlo; x <= xhi; ++x) acc += (u128d)a[x * stride] * b[(l - x) * stride];
        d[l * stride] = reduce_u128(acc, q, r0, r1, delta);
    }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// The constants travel as a __grid_constant__ kernel parameter (3.3 KB): every read is a uniform
// constant-bank access instead of a global load.
// polys x: k limbs of n, `xs` words from one polynomial to the next -> ext: nb limbs of n, `es` words apart
template <int K, int KB>
__global__ void __launch_bounds__(256) k_behz_extend(const u64 *__restrict__ x, size_t xs, u64 *__restrict__ ext, size_t es,
                                                      unsigned long long npoly, const __grid_constant__ BehzDev B)
{
    const unsigned long long total = npoly << B.logn;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = i >> B.logn;
        const uint32_t c = (uint32_t)(i & (B.n - 1));
        behz_extend_coeff_t<K, KB>(B, x + p * xs + c, B.n, ext + p * es + c, B.n);
    }
}
// dq: k limbs per polynomial (`qs` words apart), db: nb limbs per polynomial (`bs` apart) -> out: [npoly][k][n]
template <int K, int KB>
__global__ void __launch_bounds__(256) k_behz_scale(const u64 *__restrict__ dq, size_t qs, const u64 *__restrict__ db, size_t bs,
                                                     u64 *__restrict__ out, unsigned long long npoly,
                                                     const __grid_constant__ BehzDev B)
{
    const unsigned long long total = npoly << B.logn;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = i >> B.logn;
        const uint32_t c = (uint32_t)(i & (B.n - 1));
        behz_scale_coeff_t<K, KB>(B, dq + p * qs + c, B.n, db + p * bs + c, B.n, out + p * B.k * B.n + c, B.n);
    }
}

#endif  // __CUDACC__

}  // namespace fhe
