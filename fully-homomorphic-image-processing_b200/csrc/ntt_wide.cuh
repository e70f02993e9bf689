// ntt_wide.cuh -- negacyclic NTT / inverse for primes up to 61 bits (the BEHZ auxiliary base
// B_sk of Evaluator::multiply; util/smallntt over bsk_small_ntt_tables in
// SEAL 2.3, [SEAL-recall]).
//
// Same thread/pass geometry, shared-memory staging and per-pass twiddle layout as ntt.cuh (the
// load/store primitives and PassGeom are reused unchanged); only the butterfly arithmetic differs.
// The lazy mulc arithmetic of ntt.cuh needs 256 q <= 2^63, which 61-bit primes break, so these
// stages use Harvey's exact-quotient butterflies instead: values stay in [0, 4q) (forward) or
// [0, 2q) (inverse) with one conditional subtraction per butterfly, 4q < 2^63.
// Twiddles are W2 {w, floor(w 2^64 / q)} (note 2^64, not mulc's 2^63).
#pragma once
#include "ntt.cuh"

struct LimbWide {
    u64 q, q2;    // the prime (< 2^61) and 2q
    W2 ninv;      // n^-1
    W2 ninv_w1;   // n^-1 * itw[1]
    u64 r0, r1;   // floor(2^128 / q) (Barrett, for variable x variable products)
};

// x * c.w mod q in [0, 2q) for any x < 2^64 (Shoup / Harvey)
FHE_HD u64 mulsh(u64 x, W2 c, u64 q) { return x * c.w - mulhi_u64(x, c.p) * q; }

namespace nttwide {
using nttdetail::PassGeom;

template <int LOGN, int LOGE, int P, int D>
FHE_HD void fwd_stage(u64 *v, int tau, const W2 *__restrict__ tw, const LimbWide &L)
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
                const u64 x = csub(v[ia], L.q2);
                const u64 t = mulsh(v[ib], w, L.q);
                v[ia] = x + t;
                v[ib] = x - t + L.q2;
            }
        }
    }
}

template <int LOGN, int LOGE, int P, int DD>
FHE_HD void inv_stage(u64 *v, int tau, const W2 *__restrict__ itw, const LimbWide &L)
{
    using Gm = PassGeom<LOGN, LOGE, P>;
    constexpr int D = Gm::R - 1 - DD, RR = Gm::RR, half = 1 << DD;
    constexpr bool final_stage = (P == 0 && D == 0);
#pragma unroll
    for (int g = 0; g < Gm::G; ++g) {
#pragma unroll
        for (int mt = 0; mt < (1 << D); ++mt) {
            W2 w;
            if constexpr (final_stage) w = L.ninv_w1;
            else w = itw[Gm::tw_index(tau, g, D, mt)];
#pragma unroll
            for (int lo = 0; lo < half; ++lo) {
                const int M = mt * (2 * half) + lo;
                const int ia = g * RR + M, ib = ia + half;
                const u64 U = v[ia], V = v[ib];
                const u64 sum = csub(U + V, L.q2);
                const u64 dif = U - V + L.q2;
                if constexpr (final_stage) v[ia] = mulsh(sum, L.ninv, L.q);
                else v[ia] = sum;
                v[ib] = mulsh(dif, w, L.q);
            }
        }
    }
}

}  // namespace nttwide

// forward pass P on a register file: inputs in [0, 4q) (canonical on the first pass); the last pass
// leaves canonical values
template <int LOGN, int LOGE, int P>
FHE_HD void nttw_fwd_compute(u64 *v, int tau, const W2 *tw, const LimbWide &L)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int R = nttdetail::PassGeom<LOGN, LOGE, P>::R;
    nttwide::fwd_stage<LOGN, LOGE, P, 0>(v, tau, tw, L);
    if constexpr (R > 1) nttwide::fwd_stage<LOGN, LOGE, P, 1>(v, tau, tw, L);
    if constexpr (R > 2) nttwide::fwd_stage<LOGN, LOGE, P, 2>(v, tau, tw, L);
    if constexpr (R > 3) nttwide::fwd_stage<LOGN, LOGE, P, 3>(v, tau, tw, L);
    if constexpr (R > 4) nttwide::fwd_stage<LOGN, LOGE, P, 4>(v, tau, tw, L);
    if constexpr (P == S::NPASS - 1) {
#pragma unroll
        for (int e = 0; e < S::E; ++e) v[e] = csub(csub(v[e], L.q2), L.q);
    }
}
template <int LOGN, int LOGE, int P>
FHE_HD void nttw_fwd_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *tw, const LimbWide &L)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_fwd_load<LOGN, LOGE, P>(v, tau, gin, smem);
    nttw_fwd_compute<LOGN, LOGE, P>(v, tau, tw, L);
    ntt_fwd_store<LOGN, LOGE, P>(v, tau, gout, smem);
}
// inverse pass P (P = NPASS-1 first) on a register file: inputs in [0, 2q); P == 0 leaves canonical
// values times n^-1
template <int LOGN, int LOGE, int P>
FHE_HD void nttw_inv_compute(u64 *v, int tau, const W2 *itw, const LimbWide &L)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int R = nttdetail::PassGeom<LOGN, LOGE, P>::R;
    nttwide::inv_stage<LOGN, LOGE, P, 0>(v, tau, itw, L);
    if constexpr (R > 1) nttwide::inv_stage<LOGN, LOGE, P, 1>(v, tau, itw, L);
    if constexpr (R > 2) nttwide::inv_stage<LOGN, LOGE, P, 2>(v, tau, itw, L);
    if constexpr (R > 3) nttwide::inv_stage<LOGN, LOGE, P, 3>(v, tau, itw, L);
    if constexpr (R > 4) nttwide::inv_stage<LOGN, LOGE, P, 4>(v, tau, itw, L);
    if constexpr (P == 0) {
#pragma unroll
        for (int e = 0; e < S::E; ++e) v[e] = csub(v[e], L.q);
    }
}
template <int LOGN, int LOGE, int P>
FHE_HD void nttw_inv_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *itw, const LimbWide &L)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_inv_load<LOGN, LOGE, P>(v, tau, gin, smem);
    nttw_inv_compute<LOGN, LOGE, P>(v, tau, itw, L);
    ntt_inv_store<LOGN, LOGE, P>(v, tau, gout, smem);
}

#if defined(__CUDACC__)
// pass chains, as in ntt.cuh: forward passes P .. PEND-1 with a barrier after each; inverse passes P
// down to PEND with a barrier before each
template <int LOGN, int LOGE, int P, int PEND>
__device__ __forceinline__ void nttw_fwd_passes(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *tw, const LimbWide &L)
{
    if constexpr (P < PEND) {
        nttw_fwd_pass<LOGN, LOGE, P>(tau, gin, gout, smem, tw, L);
        __syncthreads();
        nttw_fwd_passes<LOGN, LOGE, P + 1, PEND>(tau, gin, gout, smem, tw, L);
    }
}
template <int LOGN, int LOGE, int P, int PEND>
__device__ __forceinline__ void nttw_inv_passes(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *itw, const LimbWide &L)
{
    if constexpr (P >= PEND) {
        __syncthreads();
        nttw_inv_pass<LOGN, LOGE, P>(tau, gin, gout, smem, itw, L);
        nttw_inv_passes<LOGN, LOGE, P - 1, PEND>(tau, gin, gout, smem, itw, L);
    }
}
#endif
