// ntt_generic.cuh -- negacyclic NTT / inverse for ANY coeff_modulus prime SEAL 2.3 accepts (up to 60
// bits, as small as 2 n + 1): the transforms of a context whose chain has a prime outside the
// [2^40, 2^55) window of the lazy arithmetic (ntt.cuh needs 256 q <= 2^63 and 32-bit quotient estimates).
//
// SEAL 2.3's own defaults (coeff_modulus_128) are 54/55-bit primes and take the lazy kernels; a chain
// with a 60-bit prime (SEAL 2.3.0's defaults, small_mods_60bit) or a small prime runs here, every limb
// of it, so that one context uses one arithmetic.  Same thread/pass geometry, shared-memory staging and
// per-pass twiddle layout as ntt.cuh (load/store primitives and PassGeom are reused); the butterflies
// are Harvey's: values in [0, 4q) (forward) or [0, 2q) (inverse), one conditional subtraction per
// butterfly, exact Shoup quotient.  Twiddles and table entries are W2 {w, floor(w 2^64 / q)} (2^64, not
// the 2^63 of mulc).  Roughly 1.7x the time of the lazy transform (profiles/r02_notes.md B).
#pragma once
#include "ntt.cuh"

// x * c.w mod q in [0, 2q) for any x < 2^64, c = {w, floor(w 2^64 / q)} (Shoup's exact quotient)
FHE_HD u64 mulsh(u64 x, W2 c, u64 q) { return x * c.w - mulhi_u64(x, c.p) * q; }

namespace nttgen {
using nttdetail::PassGeom;

template <int LOGN, int LOGE, int P, int D>
FHE_HD void fwd_stage(u64 *v, int tau, const W2 *__restrict__ tw, u64 q, u64 q2)
{
    using Gm = PassGeom<LOGN, LOGE, P>;
    constexpr int RR = Gm::RR, half = RR >> (D + 1);
#pragma unroll
    for (int g = 0; g < Gm::G; ++g) {
#pragma unroll
        for (int mt = 0; mt < (1 << D); ++mt) {
            const W2 w = tw[Gm::tw_index(tau, g, D, mt)];
#pragma unroll
            for (int lo = 0; lo < half; ++lo) {
                const int ia = g * RR + mt * (2 * half) + lo, ib = ia + half;
                const u64 x = csub(v[ia], q2);
                const u64 t = mulsh(v[ib], w, q);
                v[ia] = x + t;
                v[ib] = x - t + q2;
            }
        }
    }
}

template <int LOGN, int LOGE, int P, int DD, bool SCALE>
FHE_HD void inv_stage(u64 *v, int tau, const W2 *__restrict__ itw, const LimbDev &lp, u64 q, u64 q2)
{
    using Gm = PassGeom<LOGN, LOGE, P>;
    constexpr int D = Gm::R - 1 - DD, RR = Gm::RR, half = 1 << DD;
    constexpr bool final_stage = (P == 0 && D == 0);
#pragma unroll
    for (int g = 0; g < Gm::G; ++g) {
#pragma unroll
        for (int mt = 0; mt < (1 << D); ++mt) {
            W2 w;
            if constexpr (final_stage) w = SCALE ? lp.ninv_w1 : lp.iw1;
            else w = itw[Gm::tw_index(tau, g, D, mt)];
#pragma unroll
            for (int lo = 0; lo < half; ++lo) {
                const int M = mt * (2 * half) + lo;
                const int ia = g * RR + M, ib = ia + half;
                const u64 U = v[ia], V = v[ib];
                const u64 sum = csub(U + V, q2);
                const u64 dif = U - V + q2;
                if constexpr (final_stage && SCALE) v[ia] = mulsh(sum, lp.ninv, q);
                else v[ia] = sum;
                v[ib] = mulsh(dif, w, q);
            }
        }
    }
}
}  // namespace nttgen

// forward pass P on a register file: inputs in [0, 4q) (canonical on the first pass); the last pass
// leaves canonical values
template <int LOGN, int LOGE, int P>
FHE_HD void ntg_fwd_compute(u64 *v, int tau, const W2 *tw, const LimbDev &lp)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int R = nttdetail::PassGeom<LOGN, LOGE, P>::R;
    const u64 q = lp.a.q, q2 = 2 * q;
    nttgen::fwd_stage<LOGN, LOGE, P, 0>(v, tau, tw, q, q2);
    if constexpr (R > 1) nttgen::fwd_stage<LOGN, LOGE, P, 1>(v, tau, tw, q, q2);
    if constexpr (R > 2) nttgen::fwd_stage<LOGN, LOGE, P, 2>(v, tau, tw, q, q2);
    if constexpr (R > 3) nttgen::fwd_stage<LOGN, LOGE, P, 3>(v, tau, tw, q, q2);
    if constexpr (R > 4) nttgen::fwd_stage<LOGN, LOGE, P, 4>(v, tau, tw, q, q2);
    if constexpr (P == S::NPASS - 1) {
#pragma unroll
        for (int e = 0; e < S::E; ++e) v[e] = csub(csub(v[e], q2), q);
    }
}
template <int LOGN, int LOGE, int P>
FHE_HD void ntg_fwd_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *tw, const LimbDev &lp)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_fwd_load<LOGN, LOGE, P>(v, tau, gin, smem);
    ntg_fwd_compute<LOGN, LOGE, P>(v, tau, tw, lp);
    ntt_fwd_store<LOGN, LOGE, P>(v, tau, gout, smem);
}
// inverse pass P (P = NPASS-1 first) on a register file: inputs in [0, 2q); P == 0 leaves canonical
// values, times n^-1 when SCALE (else the preceding table carried n^-1)
template <int LOGN, int LOGE, int P, bool SCALE>
FHE_HD void ntg_inv_compute(u64 *v, int tau, const W2 *itw, const LimbDev &lp)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int R = nttdetail::PassGeom<LOGN, LOGE, P>::R;
    const u64 q = lp.a.q, q2 = 2 * q;
    nttgen::inv_stage<LOGN, LOGE, P, 0, SCALE>(v, tau, itw, lp, q, q2);
    if constexpr (R > 1) nttgen::inv_stage<LOGN, LOGE, P, 1, SCALE>(v, tau, itw, lp, q, q2);
    if constexpr (R > 2) nttgen::inv_stage<LOGN, LOGE, P, 2, SCALE>(v, tau, itw, lp, q, q2);
    if constexpr (R > 3) nttgen::inv_stage<LOGN, LOGE, P, 3, SCALE>(v, tau, itw, lp, q, q2);
    if constexpr (R > 4) nttgen::inv_stage<LOGN, LOGE, P, 4, SCALE>(v, tau, itw, lp, q, q2);
    if constexpr (P == 0) {
#pragma unroll
        for (int e = 0; e < S::E; ++e) v[e] = csub(v[e], q);
    }
}
template <int LOGN, int LOGE, int P, bool SCALE>
FHE_HD void ntg_inv_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *itw, const LimbDev &lp)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_inv_load<LOGN, LOGE, P>(v, tau, gin, smem);
    ntg_inv_compute<LOGN, LOGE, P, SCALE>(v, tau, itw, lp);
    ntt_inv_store<LOGN, LOGE, P>(v, tau, gout, smem);
}

#if defined(__CUDACC__)
template <int LOGN, int LOGE, int P, int PEND>
__device__ __forceinline__ void ntg_fwd_passes(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *tw, const LimbDev &lp)
{
    if constexpr (P < PEND) {
        ntg_fwd_pass<LOGN, LOGE, P>(tau, gin, gout, smem, tw, lp);
        __syncthreads();
        ntg_fwd_passes<LOGN, LOGE, P + 1, PEND>(tau, gin, gout, smem, tw, lp);
    }
}
template <int LOGN, int LOGE, bool SCALE, int P, int PEND>
__device__ __forceinline__ void ntg_inv_passes(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *itw, const LimbDev &lp)
{
    if constexpr (P >= PEND) {
        __syncthreads();
        ntg_inv_pass<LOGN, LOGE, P, SCALE>(tau, gin, gout, smem, itw, lp);
        ntg_inv_passes<LOGN, LOGE, SCALE, P - 1, PEND>(tau, gin, gout, smem, itw, lp);
    }
}
#endif
