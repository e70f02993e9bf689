// ntt.cuh -- negacyclic NTT / inverse NTT over one RNS limb, one CTA per limb-polynomial.
//
// B200 replacement for SEAL 2.3 util/smallntt (ntt_negacyclic_harvey, inverse_ntt_negacyclic_
// harvey; SURVEY.md 2.2 N1, Appendix A.2 -- [SEAL-recall], sources absent from /root/reference).
// Same transform (forward: Cooley-Tukey, psi powers in bit-reversed order, twiddle index m+i),
// different machine mapping:
//   * a limb-polynomial (n u64 = 32 KiB at n=4096) lives in shared memory for the whole
//     transform; each thread keeps E = 2^LOGE coefficients in registers and runs LOGE
//     butterfly stages per pass (radix-2^LOGE), so n=4096 is 3 passes / 2 shared-memory
//     exchanges instead of 12;
//   * all global traffic is coalesced: the first pass reads natural order with stride n/E,
//     the last pass writes position (e*T + tau) -- i.e. the NTT domain uses the fixed
//     permutation pi(j) = (j mod E) * T + (j div E) of SEAL's scrambled order. The NTT domain
//     is internal (every result is returned to coefficient form), so any fixed order is
//     unobservable as long as forward, tables and inverse agree;
//   * butterflies are lazy with COMPILE-TIME range tracking instead of Harvey's per-butterfly
//     conditional subtraction: forward values grow by 2q per stage, inverse sums double; the
//     slack between q and 2^64 absorbs it (class A: q < 2^57, no corrections inside a
//     forward transform; class B: q < 2^61, Harvey-like).
// Everything is __host__ __device__: tests/emu runs these exact bodies thread by thread.
#pragma once
#include "modarith.cuh"

struct LimbDev {
    u64 q;
    W2 one;      // (1, floor(2^64/q)) : Shoup "multiply by one" = lazy reduction of any u64
    W2 ninv;     // n^-1
    W2 ninv_w1;  // n^-1 * itw[1]
};

template <int LOGN, int LOGE>
struct NttShape {
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, T = N >> LOGE;
    static constexpr int NPASS = (LOGN + LOGE - 1) / LOGE;
    static constexpr int SMEM_WORDS = N + (N >> LOGE);
    FHE_HD static constexpr int s0(int p) { return p * LOGE; }
    FHE_HD static constexpr int r(int p) { return (LOGN - p * LOGE) < LOGE ? (LOGN - p * LOGE) : LOGE; }
    FHE_HD static constexpr int pad(int j) { return j + (j >> LOGE); }
};

// Range classes. FMAX: largest forward bound (units of q) a butterfly output may reach;
// LIM: largest log2-bound of an inverse butterfly operand; LBX: log2-bound at inverse pass exits.
struct RangeA { static constexpr int FMAX = 64, LIM = 6, LBX = 2; };  // q < 2^57
struct RangeB { static constexpr int FMAX = 8, LIM = 2, LBX = 2; };   // q < 2^61

namespace nttdetail {

FHE_HD constexpr int ladder_step(int b)  // smallest power of two >= ceil(b/2)
{
    int p = 1;
    while (2 * p < b) p <<= 1;
    return p;
}
// forward bound (units of q) of every value BEFORE stage s (inputs < b0*q)
template <int FMAX>
FHE_HD constexpr int fwd_bound_before(int s, int b0)
{
    int b = b0;
    for (int i = 0; i < s; ++i) {
        if (b + 2 > FMAX) b = 2;
        b += 2;
    }
    return b;
}
// inverse log2-bound of register M after `done` processed stages of a pass
template <int LIM>
FHE_HD constexpr int inv_lb(int lb_in, int M, int done)
{
    int lb = lb_in;
    for (int dd = 0; dd < done; ++dd) {
        if (lb > LIM) lb = LIM;
        lb = ((M >> dd) & 1) ? 1 : lb + 1;
    }
    return lb;
}

// v < B*q  ->  v < 2q, B compile-time
template <int B>
FHE_HD void ladder_to_2(u64 &v, u64 q)
{
    if constexpr (B > 2) {
        constexpr int p = ladder_step(B);
        v = umin_u64(v, v - (u64)p * q);
        ladder_to_2<p>(v, q);
    }
}
// v < 2^lb q -> v < 2^target q ; lb folds to a constant after unrolling
FHE_HD void ladder_log(u64 &v, int lb, int target, u64 q)
{
#pragma unroll
    for (int s = 12; s >= 1; --s)
        if (s <= lb && s > target) v = umin_u64(v, v - (q << (s - 1)));
}

template <int LOGN, int LOGE, int P>
struct PassGeom {
    using S = NttShape<LOGN, LOGE>;
    static constexpr int S0 = S::s0(P), R = S::r(P), BLO = LOGN - S0 - R, G = S::E >> R, RR = 1 << R;
    FHE_HD static int gamma(int tau, int g) { return BLO == 0 ? tau * G + g : g * S::T + tau; }
    // SEAL-order coefficient index of register e of thread tau
    FHE_HD static int index(int tau, int e)
    {
        int g = e >> R, M = e & (RR - 1), gm = gamma(tau, g);
        int H = gm >> BLO, L = gm & ((1 << BLO) - 1);
        return (H << (BLO + R)) | (M << BLO) | L;
    }
    FHE_HD static int Hof(int tau, int g) { return gamma(tau, g) >> BLO; }  // twiddle block
};

// one forward stage D of pass P on the register file v
template <int LOGN, int LOGE, int P, class RC, int D>
FHE_HD void fwd_stage(u64 *v, int tau, const W2 *__restrict__ tw, u64 q)
{
    using Gm = PassGeom<LOGN, LOGE, P>;
    constexpr int s = Gm::S0 + D, RR = Gm::RR, half = RR >> (D + 1);
    constexpr int bnd = fwd_bound_before<RC::FMAX>(s, 1);
    constexpr bool reduce = (bnd + 2 > RC::FMAX);
    const u64 two_q = 2 * q;
#pragma unroll
    for (int g = 0; g < Gm::G; ++g) {
        const int H = Gm::Hof(tau, g);
#pragma unroll
        for (int mt = 0; mt < (1 << D); ++mt) {
            const W2 w = tw[(1 << s) + (H << D) + mt];
#pragma unroll
            for (int lo = 0; lo < half; ++lo) {
                const int ia = g * RR + mt * (2 * half) + lo, ib = ia + half;
                u64 x = v[ia];
                if constexpr (reduce) ladder_to_2<bnd>(x, q);
                u64 t = shoup_lazy(v[ib], w, q);
                v[ia] = x + t;
                v[ib] = x + two_q - t;
            }
        }
    }
}

// one inverse stage (processed DD-th within pass P; undoes forward stage D = R-1-DD)
template <int LOGN, int LOGE, int P, class RC, int DD, int LB_IN>
FHE_HD void inv_stage(u64 *v, int tau, const W2 *__restrict__ itw, const LimbDev &lp)
{
    using Gm = PassGeom<LOGN, LOGE, P>;
    constexpr int D = Gm::R - 1 - DD, s = Gm::S0 + D, RR = Gm::RR, half = 1 << DD;
    constexpr bool final_stage = (P == 0 && D == 0);
    const u64 q = lp.q;
#pragma unroll
    for (int g = 0; g < Gm::G; ++g) {
        const int H = Gm::Hof(tau, g);
#pragma unroll
        for (int mt = 0; mt < (1 << D); ++mt) {
            W2 w;
            if constexpr (final_stage) w = lp.ninv_w1;
            else w = itw[(1 << s) + (H << D) + mt];
#pragma unroll
            for (int lo = 0; lo < half; ++lo) {
                const int M = mt * (2 * half) + lo;
                const int ia = g * RR + M, ib = ia + half;
                const int lb0 = inv_lb<RC::LIM>(LB_IN, M, DD);
                const int lb = lb0 > RC::LIM ? RC::LIM : lb0;
                u64 U = v[ia], V = v[ib];
                ladder_log(U, lb0, RC::LIM, q);
                ladder_log(V, lb0, RC::LIM, q);
                u64 sum = U + V;
                u64 dif = U - V + (q << lb);
                if constexpr (final_stage) v[ia] = shoup_lazy(sum, lp.ninv, q);
                else v[ia] = sum;
                v[ib] = shoup_lazy(dif, w, q);
            }
        }
    }
}

}  // namespace nttdetail

// ---------------------------------------------------------------------------------------------
// Pass primitives for thread tau.  A pass = load E registers, LOGE-or-fewer butterfly stages,
// store.  Loads/stores are split from the compute so fused kernels (NTT -> pointwise -> inverse
// NTT without leaving the SM) can chain register files directly.
//
// Forward pass P: src = global natural order (P == 0) else shared (padded); dst = shared, or
// global permuted order (e*T + tau) on the last pass.
template <int LOGN, int LOGE, int P>
FHE_HD void ntt_fwd_load(u64 *v, int tau, const u64 *gin, const u64 *smem)
{
    using S = NttShape<LOGN, LOGE>;
    using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
#pragma unroll
    for (int e = 0; e < S::E; ++e) {
        int j = Gm::index(tau, e);
        v[e] = (P == 0) ? gin[j] : smem[S::pad(j)];
    }
}
template <int LOGN, int LOGE, int P, class RC>
FHE_HD void ntt_fwd_compute(u64 *v, int tau, const W2 *__restrict__ tw, const LimbDev &lp)
{
    constexpr int R = nttdetail::PassGeom<LOGN, LOGE, P>::R;
    const u64 q = lp.q;
    nttdetail::fwd_stage<LOGN, LOGE, P, RC, 0>(v, tau, tw, q);
    if constexpr (R > 1) nttdetail::fwd_stage<LOGN, LOGE, P, RC, 1>(v, tau, tw, q);
    if constexpr (R > 2) nttdetail::fwd_stage<LOGN, LOGE, P, RC, 2>(v, tau, tw, q);
    if constexpr (R > 3) nttdetail::fwd_stage<LOGN, LOGE, P, RC, 3>(v, tau, tw, q);
    if constexpr (R > 4) nttdetail::fwd_stage<LOGN, LOGE, P, RC, 4>(v, tau, tw, q);
}
// lazy forward output (bound (1 + 2 LOGN) q in class A) -> [0, q)
template <int LOGN, int LOGE, class RC>
FHE_HD void ntt_fwd_finalize(u64 *v, const LimbDev &lp)
{
    constexpr int bout = nttdetail::fwd_bound_before<RC::FMAX>(LOGN, 1);
#pragma unroll
    for (int e = 0; e < NttShape<LOGN, LOGE>::E; ++e) {
        u64 x = v[e];
        if constexpr (bout > 4) x = shoup_lazy(x, lp.one, lp.q);  // any u64 -> [0, 2q)
        else nttdetail::ladder_to_2<bout>(x, lp.q);
        v[e] = csub(x, lp.q);
    }
}
template <int LOGN, int LOGE, int P>
FHE_HD void ntt_fwd_store(const u64 *v, int tau, u64 *gout, u64 *smem)
{
    using S = NttShape<LOGN, LOGE>;
    using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
    if constexpr (P == S::NPASS - 1) {
#pragma unroll
        for (int e = 0; e < S::E; ++e) gout[e * S::T + tau] = v[e];
    } else {
#pragma unroll
        for (int e = 0; e < S::E; ++e) smem[S::pad(Gm::index(tau, e))] = v[e];
    }
}

// Inverse pass P (run P = NPASS-1 first, P = 0 last).  src = global permuted order on the first
// executed pass (values < 2q), else shared; dst = shared, or global natural order, fully reduced
// and scaled by n^-1, on P == 0.
template <int LOGN, int LOGE, int P>
FHE_HD void ntt_inv_load(u64 *v, int tau, const u64 *gin, const u64 *smem)
{
    using S = NttShape<LOGN, LOGE>;
    using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
#pragma unroll
    for (int e = 0; e < S::E; ++e)
        v[e] = (P == S::NPASS - 1) ? gin[e * S::T + tau] : smem[S::pad(Gm::index(tau, e))];
}
template <int LOGN, int LOGE, int P, class RC>
FHE_HD void ntt_inv_compute(u64 *v, int tau, const W2 *__restrict__ itw, const LimbDev &lp)
{
    using S = NttShape<LOGN, LOGE>;
    using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
    constexpr int R = Gm::R, RR = Gm::RR;
    constexpr int LB_IN = (P == S::NPASS - 1) ? 1 : RC::LBX;
    nttdetail::inv_stage<LOGN, LOGE, P, RC, 0, LB_IN>(v, tau, itw, lp);
    if constexpr (R > 1) nttdetail::inv_stage<LOGN, LOGE, P, RC, 1, LB_IN>(v, tau, itw, lp);
    if constexpr (R > 2) nttdetail::inv_stage<LOGN, LOGE, P, RC, 2, LB_IN>(v, tau, itw, lp);
    if constexpr (R > 3) nttdetail::inv_stage<LOGN, LOGE, P, RC, 3, LB_IN>(v, tau, itw, lp);
    if constexpr (R > 4) nttdetail::inv_stage<LOGN, LOGE, P, RC, 4, LB_IN>(v, tau, itw, lp);
    if constexpr (P == 0) {
#pragma unroll
        for (int e = 0; e < S::E; ++e) v[e] = csub(v[e], lp.q);
    } else {
#pragma unroll
        for (int e = 0; e < S::E; ++e)
            nttdetail::ladder_log(v[e], nttdetail::inv_lb<RC::LIM>(LB_IN, e & (RR - 1), R), RC::LBX, lp.q);
    }
}
template <int LOGN, int LOGE, int P>
FHE_HD void ntt_inv_store(const u64 *v, int tau, u64 *gout, u64 *smem)
{
    using S = NttShape<LOGN, LOGE>;
    using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
#pragma unroll
    for (int e = 0; e < S::E; ++e) {
        int j = Gm::index(tau, e);
        if (P == 0) gout[j] = v[e];
        else smem[S::pad(j)] = v[e];
    }
}

// One whole pass (load, compute, store).  OUT_FULL: forward output reduced to [0, q).
template <int LOGN, int LOGE, int P, class RC, bool OUT_FULL>
FHE_HD void ntt_fwd_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *tw, const LimbDev &lp)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_fwd_load<LOGN, LOGE, P>(v, tau, gin, smem);
    ntt_fwd_compute<LOGN, LOGE, P, RC>(v, tau, tw, lp);
    if constexpr (OUT_FULL && P == NttShape<LOGN, LOGE>::NPASS - 1) ntt_fwd_finalize<LOGN, LOGE, RC>(v, lp);
    ntt_fwd_store<LOGN, LOGE, P>(v, tau, gout, smem);
}
template <int LOGN, int LOGE, int P, class RC>
FHE_HD void ntt_inv_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *itw, const LimbDev &lp)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_inv_load<LOGN, LOGE, P>(v, tau, gin, smem);
    ntt_inv_compute<LOGN, LOGE, P, RC>(v, tau, itw, lp);
    ntt_inv_store<LOGN, LOGE, P>(v, tau, gout, smem);
}
