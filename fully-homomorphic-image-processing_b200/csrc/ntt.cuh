// ntt.cuh -- negacyclic NTT / inverse NTT over one RNS limb, one CTA per limb-polynomial.
//
// B200 replacement for SEAL 2.3 util/smallntt (ntt_negacyclic_harvey, inverse_ntt_negacyclic_
// harvey; SURVEY.md 2.2 N1, Appendix A.2 -- [SEAL-recall], sources absent from /root/reference).
// Same transform (forward: Cooley-Tukey, psi powers in bit-reversed order, twiddle index m+i;
// inverse: Gentleman-Sande), different machine mapping:
//   * a limb-polynomial (n u64 = 32 KiB at n=4096) lives in shared memory for the whole
//     transform; each thread keeps E = 2^LOGE coefficients in registers and runs LOGE
//     butterfly stages per pass (radix-2^LOGE), so n=4096 is 3 passes / 2 shared-memory
//     exchanges instead of 12;
//   * all global traffic is coalesced: the first pass reads natural order with stride n/E,
//     the last pass writes position (e*T + tau) -- i.e. the NTT domain uses the fixed
//     permutation pi(j) = (j mod E) * T + (j div E) of SEAL's scrambled order.  The NTT domain
//     is internal (every result is returned to coefficient form), so any fixed order is
//     unobservable as long as forward, tables and inverse agree;
//   * the twiddles are stored PER PASS in the order the threads consume them
//     ([slot][thread group], PassGeom::tw_index): a warp's twiddle load is one contiguous
//     512-byte line (last pass), a two-address broadcast (middle pass) or a single broadcast
//     (first pass).  With SEAL's m+i table the last pass scattered every load over 32 lines and
//     the L1 data pipe, not arithmetic, bounded the kernel (profiles/r01_ncu_3kernel.md);
//   * butterflies are lazy with COMPILE-TIME range tracking instead of Harvey's per-butterfly
//     conditional subtraction: a forward stage adds at most 5q to a bound (modarith.cuh mulc),
//     an inverse stage doubles the bound of its sum output; bounds stay under BMAX*q < 2^63 and
//     the few registers that would exceed it at a pass boundary take a 3-instruction reduce63.
// Everything is __host__ __device__: tests/emu runs these exact bodies thread by thread.
#pragma once
#include "modarith.cuh"

struct LimbDev {
    LimbArith a;
    W2 ninv;     // n^-1
    W2 ninv_w1;  // n^-1 * itw[1]
    W2 iw1;      // itw[1] (inverse transform whose n^-1 is folded into a preceding table)
};

// coefficients per thread (2^LOGE) of the transforms of degree 2^13 and 2^14 in this build: every
// kernel, table and test takes it from here.  4: four passes, 64 registers, 1024 threads per SM;
// 5: three passes, 128 registers, 512 threads per SM.  Measured equal at n = 8192 (the fourth
// shared-memory exchange costs what the occupancy gains: multiply_plain 0.295 vs 0.302 ms per 256
// ciphertexts, multiply 2.41 vs 2.38, square 1.88 vs 1.82; profiles/r01_notes.md).  Round 2: the fused
// tensor + inverse kernel has a load-bound prologue and gains from the 4-pass shape's occupancy
// (k_tensor_inv 244 -> 231 ms per resize band, bicubic +1.4 %, decode +2.3 %; profiles/r02_notes.md E):
// n = 8192 uses 4, n = 16384 stays at 5 (a 128 KiB polynomial leaves room for one CTA either way).
#ifndef FHE_LOGE13
#define FHE_LOGE13 4
#endif
#ifndef FHE_LOGE14
#define FHE_LOGE14 5
#endif

constexpr int FHE_BMAX = 256;   // q < 2^55  =>  256 q <= 2^63
constexpr int FHE_MULC_B = 5;   // mulc output bound (units of q)
constexpr int FHE_RED_B = 4;    // reduce63 output bound

template <int LOGN, int LOGE>
struct NttShape {
    static constexpr int N = 1 << LOGN, E = 1 << LOGE, T = N >> LOGE;
    static constexpr int NPASS = (LOGN + LOGE - 1) / LOGE;
    static constexpr int SMEM_WORDS = N + (N >> LOGE);
    FHE_HD static constexpr int s0(int p) { return p * LOGE; }
    FHE_HD static constexpr int r(int p) { return (LOGN - p * LOGE) < LOGE ? (LOGN - p * LOGE) : LOGE; }
    FHE_HD static constexpr int pad(int j) { return j + (j >> LOGE); }
};

// forward output modes
enum { NTT_OUT_LAZY = 0, NTT_OUT_REDUCED = 1, NTT_OUT_CANON = 2 };

namespace nttdetail {

template <int LOGN, int LOGE, int P>
struct PassGeom {
    using S = NttShape<LOGN, LOGE>;
    static constexpr int S0 = S::s0(P), R = S::r(P), BLO = LOGN - S0 - R, G = S::E >> R, RR = 1 << R;
    static constexpr int WIDTH = S::T >> BLO;      // distinct twiddle sets among the threads
    static constexpr int SLOTS = G * (RR - 1);     // twiddles one thread consumes in this pass
    FHE_HD static constexpr int table_offset()
    {
        int off = 0;
        for (int p = 0; p < P; ++p) {
            const int r = S::r(p), blo = LOGN - S::s0(p) - r;
            off += (S::E >> r) * ((1 << r) - 1) * (S::T >> blo);
        }
        return off;
    }
    FHE_HD static int gamma(int tau, int g) { return BLO == 0 ? tau * G + g : g * S::T + tau; }
    // SEAL-order coefficient index of register e of thread tau
    FHE_HD static int index(int tau, int e)
    {
        int g = e >> R, M = e & (RR - 1), gm = gamma(tau, g);
        int H = gm >> BLO, L = gm & ((1 << BLO) - 1);
        return (H << (BLO + R)) | (M << BLO) | L;
    }
    // position in the per-pass table of the twiddle (stage D, group g, sub-block mt) of thread tau
    FHE_HD static int tw_index(int tau, int g, int D, int mt)
    {
        return table_offset() + (g * (RR - 1) + (1 << D) - 1 + mt) * WIDTH + (tau >> BLO);
    }
    // SEAL root-power table index (m + i) of the same twiddle, for thread group hq = tau >> BLO
    FHE_HD static int seal_index(int hq, int g, int D, int mt)
    {
        const int H = (BLO == 0) ? hq * G + g : hq;
        return (1 << (S0 + D)) + (H << D) + mt;
    }
};

// inverse: bound (units of q) of register M after `done` processed stages of a pass entered with
// bound b_in.  The two operands of a butterfly always have equal bounds (they differ only in the
// bit the stage processes); the difference output is a mulc result, the sum output doubles.
FHE_HD constexpr int inv_bound(int b_in, int M, int done)
{
    int b = b_in;
    for (int dd = 0; dd < done; ++dd) b = ((M >> dd) & 1) ? FHE_MULC_B : 2 * b;
    return b;
}

// one forward stage D of pass P on the register file v
template <int LOGN, int LOGE, int P, int D>
FHE_HD void fwd_stage(u64 *v, int tau, const W2 *__restrict__ tw, const LimbArith &L)
{
    using Gm = PassGeom<LOGN, LOGE, P>;
    constexpr int RR = Gm::RR, half = RR >> (D + 1);
    static_assert(1 + FHE_MULC_B * (LOGN + 1) <= FHE_BMAX, "forward lazy bound exceeds BMAX");
#pragma unroll
    for (int g = 0; g < Gm::G; ++g) {
#pragma unroll
        for (int mt = 0; mt < (1 << D); ++mt) {
            const W2 w = tw[Gm::tw_index(tau, g, D, mt)];
#pragma unroll
            for (int lo = 0; lo < half; ++lo) {
                const int ia = g * RR + mt * (2 * half) + lo, ib = ia + half;
                const u64 x = v[ia];
                const u64 t = mulc(v[ib], w, L.n2q);
                v[ia] = x + t;
                v[ib] = x + L.q5 - t;
            }
        }
    }
}

// one inverse stage (processed DD-th within pass P; undoes forward stage D = R-1-DD)
template <int LOGN, int LOGE, int P, int DD, int B_IN, bool SCALE>
FHE_HD void inv_stage(u64 *v, int tau, const W2 *__restrict__ itw, const LimbDev &lp)
{
    using Gm = PassGeom<LOGN, LOGE, P>;
    constexpr int D = Gm::R - 1 - DD, RR = Gm::RR, half = 1 << DD;
    constexpr bool final_stage = (P == 0 && D == 0);
    static_assert(2 * inv_bound(B_IN, 0, DD) <= FHE_BMAX, "inverse lazy bound exceeds BMAX");
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
                const u64 kq = bound_times_q(lp.a, inv_bound(B_IN, M, DD));  // one precomputed operand
                const u64 U = v[ia], V = v[ib];
                const u64 sum = U + V;
                const u64 dif = U - V + kq;
                if constexpr (final_stage && SCALE) v[ia] = mulc(sum, lp.ninv, lp.a.n2q);
                else v[ia] = sum;
                v[ib] = mulc(dif, w, lp.a.n2q);
            }
        }
    }
}

}  // namespace nttdetail

// bound every inverse pass is entered with (first pass: pointwise / mulc outputs; later passes:
// what the previous pass leaves after its exit reductions)
template <int LOGN, int LOGE, int P>
struct InvEntry {
    static constexpr int EXIT_MAX = FHE_BMAX >> LOGE;  // so that EXIT_MAX * 2^LOGE <= BMAX
    static constexpr int value = (P == NttShape<LOGN, LOGE>::NPASS - 1) ? FHE_MULC_B : EXIT_MAX;
    static_assert(EXIT_MAX >= FHE_MULC_B, "exit bound below the mulc output bound");
};

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
// inputs < q (canonical ciphertext / lifted plaintext); after the pass every value < (1 + 5(S0+R)) q
template <int LOGN, int LOGE, int P>
FHE_HD void ntt_fwd_compute(u64 *v, int tau, const W2 *__restrict__ tw, const LimbDev &lp)
{
    constexpr int R = nttdetail::PassGeom<LOGN, LOGE, P>::R;
    nttdetail::fwd_stage<LOGN, LOGE, P, 0>(v, tau, tw, lp.a);
    if constexpr (R > 1) nttdetail::fwd_stage<LOGN, LOGE, P, 1>(v, tau, tw, lp.a);
    if constexpr (R > 2) nttdetail::fwd_stage<LOGN, LOGE, P, 2>(v, tau, tw, lp.a);
    if constexpr (R > 3) nttdetail::fwd_stage<LOGN, LOGE, P, 3>(v, tau, tw, lp.a);
    if constexpr (R > 4) nttdetail::fwd_stage<LOGN, LOGE, P, 4>(v, tau, tw, lp.a);
}
// lazy forward output (< (1 + 5 LOGN) q) -> [0, 4q) or [0, q)
template <int LOGN, int LOGE, int OUT>
FHE_HD void ntt_fwd_finalize(u64 *v, const LimbDev &lp)
{
#pragma unroll
    for (int e = 0; e < NttShape<LOGN, LOGE>::E; ++e) {
        if constexpr (OUT == NTT_OUT_REDUCED) v[e] = reduce63(v[e], lp.a);
        if constexpr (OUT == NTT_OUT_CANON) v[e] = canon(v[e], lp.a);
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
// executed pass (values < 5q), else shared; dst = shared, or global natural order, canonical and
// scaled by n^-1 (SCALE) or assuming n^-1 was folded into the preceding table (!SCALE), on P == 0.
template <int LOGN, int LOGE, int P>
FHE_HD void ntt_inv_load(u64 *v, int tau, const u64 *gin, const u64 *smem)
{
    using S = NttShape<LOGN, LOGE>;
    using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
#pragma unroll
    for (int e = 0; e < S::E; ++e)
        v[e] = (P == S::NPASS - 1) ? gin[e * S::T + tau] : smem[S::pad(Gm::index(tau, e))];
}
template <int LOGN, int LOGE, int P, bool SCALE>
FHE_HD void ntt_inv_compute(u64 *v, int tau, const W2 *__restrict__ itw, const LimbDev &lp)
{
    using S = NttShape<LOGN, LOGE>;
    using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
    constexpr int R = Gm::R, RR = Gm::RR;
    constexpr int B_IN = InvEntry<LOGN, LOGE, P>::value;
    nttdetail::inv_stage<LOGN, LOGE, P, 0, B_IN, SCALE>(v, tau, itw, lp);
    if constexpr (R > 1) nttdetail::inv_stage<LOGN, LOGE, P, 1, B_IN, SCALE>(v, tau, itw, lp);
    if constexpr (R > 2) nttdetail::inv_stage<LOGN, LOGE, P, 2, B_IN, SCALE>(v, tau, itw, lp);
    if constexpr (R > 3) nttdetail::inv_stage<LOGN, LOGE, P, 3, B_IN, SCALE>(v, tau, itw, lp);
    if constexpr (R > 4) nttdetail::inv_stage<LOGN, LOGE, P, 4, B_IN, SCALE>(v, tau, itw, lp);
    if constexpr (P == 0) {
#pragma unroll
        for (int e = 0; e < S::E; ++e) v[e] = canon(v[e], lp.a);
    } else {
#pragma unroll
        for (int e = 0; e < S::E; ++e)
            if (nttdetail::inv_bound(B_IN, e & (RR - 1), R) > InvEntry<LOGN, LOGE, P>::EXIT_MAX)
                v[e] = reduce63(v[e], lp.a);
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

// One whole pass (load, compute, store).
template <int LOGN, int LOGE, int P, int OUT>
FHE_HD void ntt_fwd_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *tw, const LimbDev &lp)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_fwd_load<LOGN, LOGE, P>(v, tau, gin, smem);
    ntt_fwd_compute<LOGN, LOGE, P>(v, tau, tw, lp);
    if constexpr (P == NttShape<LOGN, LOGE>::NPASS - 1) ntt_fwd_finalize<LOGN, LOGE, OUT>(v, lp);
    ntt_fwd_store<LOGN, LOGE, P>(v, tau, gout, smem);
}
template <int LOGN, int LOGE, int P, bool SCALE>
FHE_HD void ntt_inv_pass(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *itw, const LimbDev &lp)
{
    u64 v[NttShape<LOGN, LOGE>::E];
    ntt_inv_load<LOGN, LOGE, P>(v, tau, gin, smem);
    ntt_inv_compute<LOGN, LOGE, P, SCALE>(v, tau, itw, lp);
    ntt_inv_store<LOGN, LOGE, P>(v, tau, gout, smem);
}

#if defined(__CUDACC__)
// Pass chains for the kernels: forward passes P .. PEND-1 with a barrier after each; inverse passes
// P down to PEND with a barrier before each.
template <int LOGN, int LOGE, int OUT, int P, int PEND>
__device__ __forceinline__ void ntt_fwd_passes(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *tw, const LimbDev &lp)
{
    if constexpr (P < PEND) {
        ntt_fwd_pass<LOGN, LOGE, P, OUT>(tau, gin, gout, smem, tw, lp);
        __syncthreads();
        ntt_fwd_passes<LOGN, LOGE, OUT, P + 1, PEND>(tau, gin, gout, smem, tw, lp);
    }
}
template <int LOGN, int LOGE, bool SCALE, int P, int PEND>
__device__ __forceinline__ void ntt_inv_passes(int tau, const u64 *gin, u64 *gout, u64 *smem, const W2 *itw, const LimbDev &lp)
{
    if constexpr (P >= PEND) {
        __syncthreads();
        ntt_inv_pass<LOGN, LOGE, P, SCALE>(tau, gin, gout, smem, itw, lp);
        ntt_inv_passes<LOGN, LOGE, SCALE, P - 1, PEND>(tau, gin, gout, smem, itw, lp);
    }
}
#endif

// Host: permute a SEAL-order root-power table (index m + i) into the per-pass layout.
template <int LOGN, int LOGE, int P = 0>
inline void ntt_permute_table(const W2 *seal_order, W2 *per_pass)
{
    using S = NttShape<LOGN, LOGE>;
    if constexpr (P < S::NPASS) {
        using Gm = nttdetail::PassGeom<LOGN, LOGE, P>;
        for (int g = 0; g < Gm::G; ++g)
            for (int D = 0; D < Gm::R; ++D)
                for (int mt = 0; mt < (1 << D); ++mt)
                    for (int hq = 0; hq < Gm::WIDTH; ++hq)
                        per_pass[Gm::tw_index(hq << Gm::BLO, g, D, mt)] = seal_order[Gm::seal_index(hq, g, D, mt)];
        ntt_permute_table<LOGN, LOGE, P + 1>(seal_order, per_pass);
    }
}
