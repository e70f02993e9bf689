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
// Data layout: a polynomial in base q is [k][n]; in base B_sk = B u {m_sk} it is [nb][n], m_sk last.
#pragma once
#include "ntt_wide.cuh"

namespace fhe {

constexpr int BEHZ_MAXK = 8;    // limbs of q supported by the multiply path
constexpr int BEHZ_MAXB = 10;   // |B_sk| <= 10

typedef unsigned __int128 u128d;

struct BehzDev {
    uint32_t n, logn, k, kb, nb;
    u64 q[BEHZ_MAXK], qr0[BEHZ_MAXK], qr1[BEHZ_MAXK];       // primes of q and floor(2^128 / q_i)
    u64 b[BEHZ_MAXB], br0[BEHZ_MAXB], br1[BEHZ_MAXB];       // B_sk
    u32 bd8[BEHZ_MAXB];                   // 8 (2^61 - b_j) = 2^64 mod b_j: the auxiliary primes are 2^61 - delta, delta < 2^26
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

struct WideRing {
    uint32_t nb;
    LimbWide limb[BEHZ_MAXB];
    const W2 *tw, *itw;  // [nb][n], per-pass layout
};

FHE_HD u64 barrett_u128(u128d z, u64 q, u64 r0, u64 r1) { return barrett128((u64)z, (u64)(z >> 64), q, r0, r1); }
// z mod p, canonical, for p = 2^61 - delta and ANY z < 2^128; d8 = 8 delta = 2^64 mod p (< 2^29).
// Two folds of the high part (2^64 = d8) and one of the top three bits (2^61 = delta): three
// 32-bit-operand multiplications instead of the five 64 x 64 products of a Barrett 128 -> 64 step.
// Every auxiliary prime of the BEHZ base has this shape (behz_host.hpp kAuxSmallMods, kMsk).
FHE_HD u64 fold61(u128d z, u64 p, u32 d8)
{
    const u64 z0 = (u64)z, z1 = (u64)(z >> 64);
    const u64 lo = (u64)(u32)z1 * d8, hi = (z1 >> 32) * d8;     // z1 * d8 = lo + hi 2^32 (< 2^93)
    const u64 t0 = lo + (hi << 32);
    u64 t1 = (hi >> 32) + (t0 < lo);
    const u64 w0 = z0 + t0;                                       // z = w0 + t1 2^64 (mod p), t1 < 2^30
    t1 += (w0 < z0);
    u64 s = w0 + t1 * d8;                                         // t1 * d8 < 2^59
    if (s < w0) s += d8;                                          // wrapped once more; now small
    const u64 r = (s & 0x1fffffffffffffffull) + (s >> 61) * (u64)(d8 >> 3);   // < 2^61 + 2^29 < 2p
    return csub(r, p);
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
            out[j * os] = fold61(acc, B.b[j], B.bd8[j]);
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
            f[j] = fold61(acc, B.b[j], B.bd8[j]);
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
    const u64 alpha = fold61(acc, msk, B.bd8[kb]);
    const bool neg = alpha > (msk >> 1);
    const u64 aabs = neg ? msk - alpha : alpha;
#pragma unroll
    for (int i = 0; i < MK; ++i)
        if (ST || i < k) {
            u128d a2 = (u128d)aabs * (neg ? B.Bq_pos[i] : B.Bq_neg[i]);
#pragma unroll
            for (int l = 0; l < MB; ++l)
                if (l < kb) a2 += (u128d)z[l] * B.I[i][l];
            out[i * os] = barrett_u128(a2, B.q[i], B.qr0[i], B.qr1[i]);
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
// words apart per poly, 128-bit accumulation (min(s1, s2) <= 32 terms of < 2^122), one Barrett reduction
FHE_HD void behz_tensor_elem(const u64 *a, const u64 *b, u64 *d, size_t stride, uint32_t s1, uint32_t s2,
                             u64 q, u64 r0, u64 r1)
{
    const uint32_t sd = s1 + s2 - 1;
    for (uint32_t l = 0; l < sd; ++l) {
        u128d acc = 0;
        const uint32_t xlo = l >= s2 ? l - (s2 - 1) : 0, xhi = l < s1 ? l : s1 - 1;
        for (uint32_t x = xlo; x <= xhi; ++x) acc += (u128d)a[x * stride] * b[(l - x) * stride];
        d[l * stride] = barrett_u128(acc, q, r0, r1);
    }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// The constants travel as a __grid_constant__ kernel parameter (3.3 KB): every read is a uniform
// constant-bank access instead of a global load.
// polys: [npoly][k][n] -> ext: [npoly][nb][n]
template <int K, int KB>
__global__ void __launch_bounds__(256) k_behz_extend(const u64 *__restrict__ x, u64 *__restrict__ ext,
                                                      unsigned long long npoly, const __grid_constant__ BehzDev B)
{
    const unsigned long long total = npoly << B.logn;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = i >> B.logn;
        const uint32_t c = (uint32_t)(i & (B.n - 1));
        behz_extend_coeff_t<K, KB>(B, x + p * B.k * B.n + c, B.n, ext + p * B.nb * B.n + c, B.n);
    }
}
// dq: [npoly][k][n], db: [npoly][nb][n] -> out: [npoly][k][n]
template <int K, int KB>
__global__ void __launch_bounds__(256) k_behz_scale(const u64 *__restrict__ dq, const u64 *__restrict__ db,
                                                     u64 *__restrict__ out, unsigned long long npoly,
                                                     const __grid_constant__ BehzDev B)
{
    const unsigned long long total = npoly << B.logn;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = i >> B.logn;
        const uint32_t c = (uint32_t)(i & (B.n - 1));
        behz_scale_coeff_t<K, KB>(B, dq + p * B.k * B.n + c, B.n, db + p * B.nb * B.n + c, B.n, out + p * B.k * B.n + c, B.n);
    }
}

// Tensor product in the NTT domain, one modulus set (nl limbs): for every ciphertext pair and
// every output poly l:  D[l] = sum_{x + y = l} A[x] (.) B[y]  per limb and position, 128-bit
// accumulation, one Barrett reduction.  A: [n_ct][s1][nl][n], Bm: [n_ct][s2][nl][n] (may be A),
// D: [n_ct][s1 + s2 - 1][nl][n].  mod / r0 / r1: [nl].
__global__ void __launch_bounds__(256) k_behz_tensor(const u64 *__restrict__ A, const u64 *__restrict__ Bm,
                                                      u64 *__restrict__ D, uint32_t s1, uint32_t s2, uint32_t nl,
                                                      uint32_t logn, unsigned long long n_ct,
                                                      const u64 *__restrict__ mod, const u64 *__restrict__ r0,
                                                      const u64 *__restrict__ r1)
{
    // grid = (nl * n / 256, ciphertexts (strided)): r = limb * n + pos without a 64-bit division
    const uint32_t sd = s1 + s2 - 1;
    const unsigned long long per_ct = (unsigned long long)nl << logn;
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = (uint32_t)(r >> logn);
    const u64 m = mod[j], m0 = r0[j], m1 = r1[j];
    for (unsigned long long ct = blockIdx.y; ct < n_ct; ct += gridDim.y) {
        const u64 *a = A + ct * s1 * per_ct + r, *b = Bm + ct * s2 * per_ct + r;
        u64 *d = D + ct * sd * per_ct + r;
        behz_tensor_elem(a, b, d, per_ct, s1, s2, m, m0, m1);
    }
}

// NTT over B_sk: one CTA per limb-polynomial, limb = blockIdx % nb
template <int LOGN, int LOGE>
__global__ void __launch_bounds__(1 << (LOGN - LOGE)) k_nttw_fwd(const u64 *in, u64 *out, WideRing R)
{
    using S = NttShape<LOGN, LOGE>;
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.nb), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *tw = R.tw + (size_t)j * S::N;
    const LimbWide lp = R.limb[j];
    nttw_fwd_pass<LOGN, LOGE, 0>(tau, gin, gout, smem, tw, lp);
    __syncthreads();
    nttw_fwd_pass<LOGN, LOGE, 1>(tau, gin, gout, smem, tw, lp);
    __syncthreads();
    nttw_fwd_pass<LOGN, LOGE, 2>(tau, gin, gout, smem, tw, lp);
    if constexpr (S::NPASS == 4) {
        __syncthreads();
        nttw_fwd_pass<LOGN, LOGE, 3>(tau, gin, gout, smem, tw, lp);
    }
    static_assert(S::NPASS == 3 || S::NPASS == 4, "3 or 4 passes");
}
template <int LOGN, int LOGE>
__global__ void __launch_bounds__(1 << (LOGN - LOGE)) k_nttw_inv(const u64 *in, u64 *out, WideRing R)
{
    using S = NttShape<LOGN, LOGE>;
    static_assert(S::NPASS == 3 || S::NPASS == 4, "3 or 4 passes");
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.nb), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *itw = R.itw + (size_t)j * S::N;
    const LimbWide lp = R.limb[j];
    if constexpr (S::NPASS == 4) {
        nttw_inv_pass<LOGN, LOGE, 3>(tau, gin, gout, smem, itw, lp);
        __syncthreads();
    }
    nttw_inv_pass<LOGN, LOGE, 2>(tau, gin, gout, smem, itw, lp);
    __syncthreads();
    nttw_inv_pass<LOGN, LOGE, 1>(tau, gin, gout, smem, itw, lp);
    __syncthreads();
    nttw_inv_pass<LOGN, LOGE, 0>(tau, gin, gout, smem, itw, lp);
}
#endif  // __CUDACC__

}  // namespace fhe
