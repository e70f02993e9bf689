// fhe_b200.cu -- kernels and C-ABI of libfhe_b200.so (see include/fhe_b200.h).
//
// sm_100a only.  No tensor cores: the path is 64-bit integer modular arithmetic on the
// IMAD/ALU pipes, staged through shared memory, with coalesced HBM access (north_star).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fhe_b200.h"
#include "behz_host.hpp"
#include "dct.cuh"
#include "ntt.cuh"
#include "ntt_generic.cuh"
#include "ring_host.hpp"

using namespace fhe;

// =============================================================================================
// device-side context
// =============================================================================================
struct RingDev {
    uint32_t n, logn, k;
    u64 t, plain_thr, q_mod_t;
    LimbDev limb[16];
    u64 delta[16], ratio0[16], ratio1[16];
    const W2 *tw, *itw;  // [k][n]
};

// =============================================================================================
// NTT kernels: one CTA per limb-polynomial, n/E threads (csrc/ntt.cuh)
// =============================================================================================
// resident CTAs per SM the register allocation is held to: 16 coefficients per thread fit 64
// registers (1024 threads per SM), 32 coefficients per thread need 128 (512 threads per SM)
#ifndef FHE_NTT_OCC_THREADS
#define FHE_NTT_OCC_THREADS 1024
#endif
#define NTT_MIN_CTAS(LOGN, LOGE) ((((LOGE) == 4 ? FHE_NTT_OCC_THREADS : 512) >> ((LOGN) - (LOGE))) > 0 ? (((LOGE) == 4 ? FHE_NTT_OCC_THREADS : 512) >> ((LOGN) - (LOGE))) : 1)
template <int LOGN, int LOGE, int OUT>
__global__ void __launch_bounds__(1 << (LOGN - LOGE), NTT_MIN_CTAS(LOGN, LOGE)) k_ntt_fwd(const u64 *in, u64 *out, RingDev R)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int LAST = S::NPASS - 1;
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.k), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *tw = R.tw + (size_t)j * S::N;
    const LimbDev lp = R.limb[j];
    ntt_fwd_passes<LOGN, LOGE, OUT, 0, LAST>(tau, gin, gout, smem, tw, lp);
    ntt_fwd_pass<LOGN, LOGE, LAST, OUT>(tau, gin, gout, smem, tw, lp);
}

template <int LOGN, int LOGE, bool SCALE>
__global__ void __launch_bounds__(1 << (LOGN - LOGE), NTT_MIN_CTAS(LOGN, LOGE)) k_ntt_inv(const u64 *in, u64 *out, RingDev R)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int LAST = S::NPASS - 1;
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.k), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *itw = R.itw + (size_t)j * S::N;
    const LimbDev lp = R.limb[j];
    ntt_inv_pass<LOGN, LOGE, LAST, SCALE>(tau, gin, gout, smem, itw, lp);
    ntt_inv_passes<LOGN, LOGE, SCALE, LAST - 1, 0>(tau, gin, gout, smem, itw, lp);
}

// The same three kernels for a context with a prime outside the window of the lazy arithmetic
// (ntt_generic.cuh: Harvey butterflies, any prime below 2^61): forward (canonical output), inverse,
// multiply_plain.
template <int LOGN, int LOGE>
__global__ void __launch_bounds__(1 << (LOGN - LOGE)) k_ntt_fwd_g(const u64 *in, u64 *out, RingDev R)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int LAST = S::NPASS - 1;
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.k), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *tw = R.tw + (size_t)j * S::N;
    const LimbDev lp = R.limb[j];
    ntg_fwd_passes<LOGN, LOGE, 0, LAST>(tau, gin, gout, smem, tw, lp);
    ntg_fwd_pass<LOGN, LOGE, LAST>(tau, gin, gout, smem, tw, lp);
}
template <int LOGN, int LOGE, bool SCALE>
__global__ void __launch_bounds__(1 << (LOGN - LOGE)) k_ntt_inv_g(const u64 *in, u64 *out, RingDev R)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int LAST = S::NPASS - 1;
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.k), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *itw = R.itw + (size_t)j * S::N;
    const LimbDev lp = R.limb[j];
    ntg_inv_pass<LOGN, LOGE, LAST, SCALE>(tau, gin, gout, smem, itw, lp);
    ntg_inv_passes<LOGN, LOGE, SCALE, LAST - 1, 0>(tau, gin, gout, smem, itw, lp);
}
template <int LOGN, int LOGE>
__global__ void __launch_bounds__(1 << (LOGN - LOGE))
k_mulplain_g(const u64 *in, u64 *out, const W2 *__restrict__ table, RingDev R)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int LAST = S::NPASS - 1;
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.k), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *tw = R.tw + (size_t)j * S::N, *itw = R.itw + (size_t)j * S::N;
    const W2 *tab = table + (size_t)j * S::N;
    const LimbDev lp = R.limb[j];
    ntg_fwd_passes<LOGN, LOGE, 0, LAST>(tau, gin, gout, smem, tw, lp);
    u64 v[S::E];
    ntt_fwd_load<LOGN, LOGE, LAST>(v, tau, gin, smem);
    ntg_fwd_compute<LOGN, LOGE, LAST>(v, tau, tw, lp);
#pragma unroll
    for (int e = 0; e < S::E; ++e) v[e] = mulsh(v[e], tab[e * S::T + tau], lp.a.q);  // [0, 2q): the inverse's input range
    ntg_inv_compute<LOGN, LOGE, LAST, false>(v, tau, itw, lp);
    ntt_inv_store<LOGN, LOGE, LAST>(v, tau, gout, smem);
    ntg_inv_passes<LOGN, LOGE, false, LAST - 1, 0>(tau, gin, gout, smem, itw, lp);
}

// Evaluator::multiply_plain as ONE kernel per limb-poly: forward NTT (lazy), dyadic product with
// the pre-transformed plaintext (times n^-1), inverse NTT -- the polynomial is read from HBM once
// and written once; the NTT-domain values never leave the SM.
template <int LOGN, int LOGE>
__global__ void __launch_bounds__(1 << (LOGN - LOGE))
k_mulplain(const u64 *in, u64 *out, const W2 *__restrict__ table, RingDev R)
{
    using S = NttShape<LOGN, LOGE>;
    constexpr int LAST = S::NPASS - 1;
    extern __shared__ u64 smem[];
    const size_t p = blockIdx.x;
    const int j = (int)(p % R.k), tau = threadIdx.x;
    const u64 *gin = in + p * S::N;
    u64 *gout = out + p * S::N;
    const W2 *tw = R.tw + (size_t)j * S::N, *itw = R.itw + (size_t)j * S::N;
    const W2 *tab = table + (size_t)j * S::N;
    const LimbDev lp = R.limb[j];
    ntt_fwd_passes<LOGN, LOGE, NTT_OUT_LAZY, 0, LAST>(tau, gin, gout, smem, tw, lp);
    u64 v[S::E];
    ntt_fwd_load<LOGN, LOGE, LAST>(v, tau, gin, smem);
    ntt_fwd_compute<LOGN, LOGE, LAST>(v, tau, tw, lp);
#pragma unroll
    for (int e = 0; e < S::E; ++e) v[e] = mulc(v[e], tab[e * S::T + tau], lp.a.n2q);
    ntt_inv_compute<LOGN, LOGE, LAST, false>(v, tau, itw, lp);
    ntt_inv_store<LOGN, LOGE, LAST>(v, tau, gout, smem);
    ntt_inv_passes<LOGN, LOGE, false, LAST - 1, 0>(tau, gin, gout, smem, itw, lp);
}

// Evaluator::multiply_plain by a plaintext with a few +-1 coefficients: the ring product is a sum
// of signed negacyclic rotations, out[i] = sum_t s_t * in[i - e_t] (sign flipped when the index
// wraps), exact like the NTT route and therefore the same canonical residues -- but one read and
// one write of the polynomial and no multiplication at all.  One CTA per limb-polynomial, staged in
// shared memory so the operation is in place.
constexpr int SPARSE_MAX = 8;
struct SparseTerms {
    uint32_t n_terms;
    uint32_t term[SPARSE_MAX];  // exponent | sign << 31
};
__global__ void __launch_bounds__(256) k_mulplain_sparse(u64 *data, SparseTerms T, RingDev R)
{
    extern __shared__ u64 spoly[];
    const uint32_t n = R.n;
    const size_t p = blockIdx.x;
    const u64 q = R.limb[p % R.k].a.q;
    u64 *g = data + p * n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) spoly[i] = g[i];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        u64 acc = 0;  // < (n_terms + 1) q: every term adds a value in [0, q)
        for (uint32_t t = 0; t < T.n_terms; ++t) {
            const uint32_t e = T.term[t] & 0x7fffffffu;
            const bool neg = (T.term[t] >> 31) != 0;
            const uint32_t src = (i - e) & (n - 1);
            const u64 v = spoly[src];
            const bool flip = neg != (i < e);  // x^n = -1
            acc += flip ? negmod(v, q) : v;
        }
        while (acc >= q) acc -= q;  // at most SPARSE_MAX - 1 subtractions
        g[i] = acc;
    }
}

// =============================================================================================
// pointwise DCT (+quantize) kernel (csrc/dct.cuh)
// =============================================================================================
// CTA = 8 warps on one (limb, 32-position tile); loops over `upc` consecutive units
// (unit = one poly of one block: 64 limb-polys, one per ciphertext of the block) so the
// 12 + 8 constants a thread needs stay in registers across units (128 registers, 2 CTAs per SM).
// Measured alternatives (profiles/r01_notes.md): per-thread LDG/STG of the unit instead of TMA
// 43.3 ms per 8192 blocks against 34.3; constants in shared memory at 1024 threads per SM 45.7;
// 64 row-sized cp.async.bulk per unit instead of one tensor box 57.0.
//
// The ciphertext array [ct][poly*k+limb][n] is described to the TMA unit as a
// rank-3 u64 tensor; a unit's 64 x 32 inputs (row i = 32 consecutive NTT positions of ciphertext i
// of the block) are ONE box {32, 1, 64}, fetched by one cp.async.bulk.tensor (SASS UTMALDG) into a
// 16 KiB shared-memory stage and written back by one UTMASTG.  The stage is input buffer, 8x8
// transposition buffer and output buffer at once (the row phase reads and writes rows 8r..8r+7 in
// place, the column phase rows 8i+col).  With a ring of DCT_NSTAGE stages the loads of the next
// units and the stores of the previous ones run in the async proxy under the arithmetic; the
// threads issue no global loads or stores at all.  Ordering: one thread issues every copy, so its
// own cp.async.bulk.wait_group.read is the WAR guard between a stage's store and the next load
// into it; load completion is an mbarrier transaction count.
constexpr int DCT_NSTAGE = 3;
constexpr uint32_t DCT_UNIT_BYTES = 64 * DCT_TILE * sizeof(u64);
constexpr size_t DCT_SMEM = (size_t)DCT_NSTAGE * DCT_UNIT_BYTES + 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(256, 2)
k_dct_pointwise(const __grid_constant__ CUtensorMap tmap, const W2 *__restrict__ ctab, const W2 *__restrict__ stab,
                    RingDev R, unsigned long long nunits, int upc, uint32_t blk_stride)
{
    extern __shared__ __align__(128) unsigned char dsm_raw[];
    u64 *stage = reinterpret_cast<u64 *>(dsm_raw);                         // [NSTAGE][64][TILE]
    const uint32_t bars = smem_u32(dsm_raw + (size_t)DCT_NSTAGE * DCT_UNIT_BYTES);  // NSTAGE mbarriers
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const uint32_t n = R.n, k = R.k, tiles = n / DCT_TILE;
    const uint32_t tile = blockIdx.x % tiles, j = (blockIdx.x / tiles) % k;
    const unsigned long long ug = blockIdx.x / (tiles * k);
    const size_t pos = (size_t)tile * DCT_TILE + lane;
    const u64 q = R.limb[j].a.q, n2q = R.limb[j].a.n2q;
    const unsigned long long u0 = ug * upc;
    const int nu = (int)(nunits - u0 < (unsigned long long)upc ? nunits - u0 : upc);
    // box coordinates of unit u0 + it: (position, poly * k + limb, first ciphertext of the block)
    const int c0 = (int)(tile * DCT_TILE);
    auto load_unit = [&](int it) {
        const unsigned long long u = u0 + it;
        const int b = it % DCT_NSTAGE, c1 = (int)((u & 1) * k + j), c2 = (int)((u >> 1) * blk_stride);
        const uint32_t bar = bars + 8 * b, dst = smem_u32(stage + (size_t)b * 64 * DCT_TILE);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(DCT_UNIT_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(dst), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
    };
    if (tid == 0) {
        for (int b = 0; b < DCT_NSTAGE; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap));
    }
    __syncthreads();
    if (tid == 0)
        for (int it = 0; it < DCT_NSTAGE - 1 && it < nu; ++it) load_unit(it);
    W2 c[DCT_NCONST], s[8];
#pragma unroll
    for (int i = 0; i < DCT_NCONST; ++i) c[i] = ctab[((size_t)i * k + j) * n + pos];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = stab[((size_t)(8 * i + w) * k + j) * n + pos];
    for (int it = 0; it < nu; ++it) {
        const int b = it % DCT_NSTAGE;
        u64 *sb = stage + (size_t)b * 64 * DCT_TILE;
        mbar_wait(bars + 8 * b, (uint32_t)(it / DCT_NSTAGE) & 1);
        dct_row_phase(w, lane, sb, DCT_TILE, sb, c, q, n2q);
        __syncthreads();
        dct_col_phase(w, lane, sb, DCT_TILE, sb, c, s, 1, q, n2q);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA store
        __syncthreads();
        if (tid == 0) {
            const unsigned long long u = u0 + it;
            const int c1 = (int)((u & 1) * k + j), c2 = (int)((u >> 1) * blk_stride);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(sb)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the stage of unit it-1 is the one unit it+NSTAGE-1 loads into: its store (the previous
            // bulk group of this thread) must have finished reading shared memory
            if (it + DCT_NSTAGE - 1 < nu) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                load_unit(it + DCT_NSTAGE - 1);
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores done before the CTA's smem is released
}

static int fail(int code, const std::string &msg);
// Tensor map of a ciphertext array for the kernel above (host side; the driver entry point is
// resolved through the runtime so the library does not link libcuda directly).
static int make_ct_tensor_map(CUtensorMap *tm, const u64 *base, uint64_t n_ct, uint32_t k, uint32_t n)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
            return fail(FHE_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        encode = (encode_fn)fn;
    }
    const cuuint64_t dims[3] = {n, 2ull * k, n_ct};
    const cuuint64_t strides[2] = {(cuuint64_t)n * 8, (cuuint64_t)2 * k * n * 8};
    const cuuint32_t box[3] = {(cuuint32_t)DCT_TILE, 1, 64}, estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void *)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FHE_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return FHE_OK;
}

// =============================================================================================
// element-wise kernels
// =============================================================================================
// flat arrays of limb-polys [..][limb][n]; limb of element i is (i / n) % k
template <int OP>  // 0 add, 1 sub, 2 negate
__global__ void k_elementwise(u64 *a, const u64 *b, unsigned long long total, RingDev R)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const u64 q = R.limb[(i >> R.logn) % R.k].a.q;
        u64 x = a[i];
        if (OP == 0) x = addmod(x, b[i], q);
        else if (OP == 1) x = submod(x, b[i], q);
        else x = negmod(x, q);
        a[i] = x;
    }
}

// centred lift of plaintexts [count][n] (< t) into [count][k][n]  
__global__ void k_plain_lift(const u64 *plain, u64 *out, unsigned long long total, RingDev R)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i & (R.n - 1)), j = (uint32_t)((i >> R.logn) % R.k);
        const unsigned long long pl = (i >> R.logn) / R.k;
        const u64 m = plain[pl * R.n + c];
        out[i] = m >= R.plain_thr ? m + (R.limb[j].a.q - R.t) : m;
    }
}
// Delta-scaling of plaintexts (Encryptor::preencrypt): [count][n] -> [count][k][n]
__global__ void k_plain_scale(const u64 *plain, u64 *out, unsigned long long total, RingDev R)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i & (R.n - 1)), j = (uint32_t)((i >> R.logn) % R.k);
        const unsigned long long pl = (i >> R.logn) / R.k;
        const u64 m = plain[pl * R.n + c];
        u64 lo = R.delta[j] * m, hi = mulhi_u64(R.delta[j], m);
        if (m >= R.plain_thr) {
            u64 s = lo + R.q_mod_t;
            hi += (s < lo);
            lo = s;
        }
        out[i] = barrett128(lo, hi, R.limb[j].a.q, R.ratio0[j], R.ratio1[j]);
    }
}
// w -> (w, floor(w 2^63 / q)), the mulc constant form; optional second factor (w = a*b mod q
// first) and optional factor n^-1 (tables that feed an inverse NTT without its own scaling)
__global__ void k_make_shoup(const u64 *a, const u64 *b, W2 *out, unsigned long long total,
                             unsigned long long b_period, int times_ninv, RingDev R, int scale_bits = 63)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)((i >> R.logn) % R.k);
        const u64 q = R.limb[j].a.q;
        u64 w = a[i];
        if (b) w = mulmod_barrett(w, b[i % b_period], q, R.ratio0[j], R.ratio1[j]);
        if (times_ninv) w = mulmod_barrett(w, R.limb[j].ninv.w, q, R.ratio0[j], R.ratio1[j]);
        W2 r;
        r.w = w;
        r.p = (u64)((((unsigned __int128)w) << scale_bits) / q);  // 63: mulc; 64: mulsh (generic contexts)
        out[i] = r;
    }
}
// poly 0 of every ciphertext +/- scaled plaintext
template <int SIGN>
__global__ void k_plain_addsub(u64 *a, const u64 *scaled /*[k][n]*/, unsigned long long n_ct,
                               uint32_t size, RingDev R)
{
    const unsigned long long per = (unsigned long long)R.k * R.n, total = n_ct * per;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long ct = i / per, r = i % per;
        const u64 q = R.limb[r >> R.logn].a.q;
        u64 *p = a + ct * per * size + r;
        *p = SIGN > 0 ? addmod(*p, scaled[r], q) : submod(*p, scaled[r], q);
    }
}

// element i of dst <- src (one ciphertext of `words` words), i < count
__global__ void k_broadcast(u64 *dst, const u64 *__restrict__ src, unsigned long long words, unsigned long long count)
{
    const unsigned long long total = words * count;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = src[i % words];
}
// poly 0 of ciphertext i += Delta-scaling of its own short plaintext plains[i][0 .. ncoef) (the other
// coefficients of the plaintext are zero); as k_plain_scale + k_plain_addsub, per element
__global__ void k_plain_add_each(u64 *a, const u64 *__restrict__ plains, unsigned long long n_ct, uint32_t size, uint32_t ncoef, RingDev R)
{
    const unsigned long long total = n_ct * R.k * ncoef;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i % ncoef);
        const uint32_t j = (uint32_t)((i / ncoef) % R.k);
        const unsigned long long ct = i / ((unsigned long long)ncoef * R.k);
        const u64 m = plains[ct * ncoef + c];
        if (!m) continue;
        u64 lo = R.delta[j] * m, hi = mulhi_u64(R.delta[j], m);
        if (m >= R.plain_thr) {
            u64 s = lo + R.q_mod_t;
            hi += (s < lo);
            lo = s;
        }
        const u64 q = R.limb[j].a.q;
        u64 *p = a + (ct * size * R.k + j) * R.n + c;
        *p = addmod(*p, barrett128(lo, hi, q, R.ratio0[j], R.ratio1[j]), q);
    }
}

// synthetic inputs: same generator as oracle/fhe_oracle.c orc_fill_uniform
__global__ void k_fill_uniform(u64 *dst, unsigned long long first_ct, unsigned long long total,
                               uint32_t size, u64 seed, RingDev R)
{
    const unsigned long long per_ct = (unsigned long long)size * R.k * R.n;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const u64 g = first_ct * per_ct + i;  // global element index
        const u64 q = R.limb[(g >> R.logn) % R.k].a.q;
        dst[i] = mulhi_u64(splitmix64(seed + g * 0x9E3779B97F4A7C15ull), q);
    }
}

__global__ void k_checksum(const u64 *v, unsigned long long count, unsigned long long *out)
{
    u64 acc = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count;
         i += (unsigned long long)gridDim.x * blockDim.x)
        acc += splitmix64(v[i] + i * 0x9E3779B97F4A7C15ull);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// =============================================================================================
// host side
// =============================================================================================
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fail(FHE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    } while (0)

struct fhe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    bool own_stream = false;
    HostRing host;
    RingDev dev{};
    W2 *d_tw = nullptr, *d_itw = nullptr;
    int sm_count = 148;
    bool generic = false;  // a coeff_modulus prime outside [2^40, 2^55): ntt_generic.cuh kernels, no fused DCT
    uint64_t launches = 0;
    uint32_t dct_chunk_blocks = 0;
    // staging for the host-buffer path
    static constexpr int kSlots = 3;
    u64 *stage[kSlots] = {nullptr, nullptr, nullptr};
    size_t stage_blocks = 0;
    cudaEvent_t ev_in[kSlots]{}, ev_comp[kSlots]{}, ev_out[kSlots]{};
    // fhe_upload_async / fhe_download_async / fhe_transfer_fence
    static constexpr int kDown = 8;
    cudaEvent_t ev_mark = nullptr, ev_up = nullptr, ev_down[kDown]{}, ev_upr[kDown]{};
    bool up_pending = false;
    uint64_t down_issued = 0, down_waited = 0;  // downloads issued / known complete
    uint64_t up_issued = 0, up_waited = 0;      // uploads issued / known complete (fhe_upload_fence)
    unsigned long long *d_scratch = nullptr;
    // optional per-kernel timing of the DCT path
    bool stage_timing = false;
    std::vector<cudaEvent_t> stage_ev;  // 4 per chunk of the last call
    // exact-size cache in front of the stream-ordered allocator (fhe_device_alloc / fhe_device_free):
    // every buffer is used on this context's one stream only, so a freed block may be handed out
    // again at once, and the seal-shaped host layer asks for the same few sizes over and over
    std::unordered_map<void *, size_t> alloc_size;
    std::unordered_map<size_t, std::vector<void *>> alloc_free;
    size_t alloc_cached = 0, alloc_cache_limit = 0;
    // pinned ring for small host -> device parameter uploads that must not synchronise the stream
    // (stage_params): kRing slots of kStageSlot bytes of MAPPED pinned memory, one event per slot
    static constexpr int kRing = 64;
    static constexpr size_t kStageSlot = 64 << 10;
    char *h_ring = nullptr, *d_ring = nullptr;
    cudaEvent_t ev_ring[kRing]{};
    bool ring_used[kRing]{};
    int ring_next = 0;
    // per-kernel device time (fhe_profile_begin / fhe_profile_end): an event after every launch
    bool prof_on = false;
    std::vector<std::pair<const char *, cudaEvent_t>> prof_marks;
    // Evaluator::multiply / square / relinearize state, built on first use (behz_api.inc)
    struct BehzState *behz = nullptr;
};
static void behz_state_destroy(struct BehzState *s);
struct fhe_ptab {
    fhe_ctx *ctx = nullptr;
    uint32_t count = 0;
    W2 *d_ntt = nullptr;     // [count][k][n] Shoup pairs of NTT(lift(plain))
    u64 *d_scaled = nullptr; // [count][k][n] Delta-scaled coefficients
    std::vector<uint8_t> is_zero;
    // plaintexts whose non-zero coefficients are all +1 or -1 (t - 1) and at most SPARSE_MAX many:
    // exponent | sign << 31 per term; empty = not sparse.  FractionalEncoder encodings of small
    // integers and of powers of two are like that (3 = 1 + x, 5 = 1 + x^2, 0.5 = -x^(n-1), 0.125 = -x^(n-3)).
    std::vector<std::vector<uint32_t>> sparse;
};
struct fhe_dct_plan {
    fhe_ctx *ctx = nullptr;
    W2 *d_c = nullptr;  // [12][k][n]
    W2 *d_s = nullptr;  // [64][k][n]
    bool has_quant = false;
};

// every kernel launch of the library passes here
static inline void launched(fhe_ctx *c, const char *kernel)
{
    c->launches++;
    if (c->prof_on) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) == cudaSuccess) {
            cudaEventRecord(e, c->stream);
            c->prof_marks.emplace_back(kernel, e);
        }
    }
}

static void clear_stage_events(fhe_ctx *c)
{
    for (cudaEvent_t e : c->stage_ev) cudaEventDestroy(e);
    c->stage_ev.clear();
}

static inline int grid_for(unsigned long long total, int threads, int sm_count)
{
    unsigned long long g = (total + threads - 1) / threads;
    unsigned long long cap = (unsigned long long)sm_count * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

template <class K>
static int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FHE_OK;
}

// dispatch on ring degree: (LOGN, LOGE) pairs with exactly three passes
#define FHE_DISPATCH_LOGN(logn, CALL)                              \
    switch (logn) {                                                \
    case 10: { CALL(10, 4); } break;                               \
    case 11: { CALL(11, 4); } break;                               \
    case 12: { CALL(12, 4); } break;                               \
    case 13: { CALL(13, FHE_LOGE13); } break;                               \
    case 14: { CALL(14, FHE_LOGE14); } break;                               \
    default: return fail(FHE_ERR_UNSUPPORTED, "unsupported ring degree"); \
    }

template <int LOGN, int LOGE, int OUT>
static int launch_fwd_mode(fhe_ctx *c, const u64 *in, u64 *out, uint64_t npoly, cudaStream_t st)
{
    using S = NttShape<LOGN, LOGE>;
    const size_t smem = S::SMEM_WORDS * sizeof(u64);
    if (int rc = set_smem(k_ntt_fwd<LOGN, LOGE, OUT>, smem)) return rc;
    k_ntt_fwd<LOGN, LOGE, OUT><<<(unsigned)npoly, S::T, smem, st>>>(in, out, c->dev);
    launched(c, "k_ntt_fwd");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}
template <int LOGN, int LOGE>
static int launch_fwd_generic(fhe_ctx *c, const u64 *in, u64 *out, uint64_t npoly, cudaStream_t st)
{
    using S = NttShape<LOGN, LOGE>;
    const size_t smem = S::SMEM_WORDS * sizeof(u64);
    if (int rc = set_smem(k_ntt_fwd_g<LOGN, LOGE>, smem)) return rc;
    k_ntt_fwd_g<LOGN, LOGE><<<(unsigned)npoly, S::T, smem, st>>>(in, out, c->dev);
    launched(c, "k_ntt_fwd_g");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}
template <int LOGN, int LOGE, bool SCALE>
static int launch_inv_generic(fhe_ctx *c, const u64 *in, u64 *out, uint64_t npoly, cudaStream_t st)
{
    using S = NttShape<LOGN, LOGE>;
    const size_t smem = S::SMEM_WORDS * sizeof(u64);
    if (int rc = set_smem(k_ntt_inv_g<LOGN, LOGE, SCALE>, smem)) return rc;
    k_ntt_inv_g<LOGN, LOGE, SCALE><<<(unsigned)npoly, S::T, smem, st>>>(in, out, c->dev);
    launched(c, "k_ntt_inv_g");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}
template <int LOGN, int LOGE>
static int launch_fwd(fhe_ctx *c, const u64 *in, u64 *out, uint64_t npoly, int mode, cudaStream_t st)
{
    if (c->generic) return launch_fwd_generic<LOGN, LOGE>(c, in, out, npoly, st);  // canonical in every mode
    if (mode == NTT_OUT_CANON) return launch_fwd_mode<LOGN, LOGE, NTT_OUT_CANON>(c, in, out, npoly, st);
    return launch_fwd_mode<LOGN, LOGE, NTT_OUT_REDUCED>(c, in, out, npoly, st);
}
template <int LOGN, int LOGE, bool SCALE>
static int launch_inv(fhe_ctx *c, const u64 *in, u64 *out, uint64_t npoly, cudaStream_t st)
{
    if (c->generic) return launch_inv_generic<LOGN, LOGE, SCALE>(c, in, out, npoly, st);
    using S = NttShape<LOGN, LOGE>;
    const size_t smem = S::SMEM_WORDS * sizeof(u64);
    if (int rc = set_smem(k_ntt_inv<LOGN, LOGE, SCALE>, smem)) return rc;
    k_ntt_inv<LOGN, LOGE, SCALE><<<(unsigned)npoly, S::T, smem, st>>>(in, out, c->dev);
    launched(c, "k_ntt_inv");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}
template <int LOGN, int LOGE>
static int launch_mulplain(fhe_ctx *c, const u64 *in, u64 *out, const W2 *tab, uint64_t npoly, cudaStream_t st)
{
    using S = NttShape<LOGN, LOGE>;
    const size_t smem = S::SMEM_WORDS * sizeof(u64);
    if (c->generic) {
        if (int rc = set_smem(k_mulplain_g<LOGN, LOGE>, smem)) return rc;
        k_mulplain_g<LOGN, LOGE><<<(unsigned)npoly, S::T, smem, st>>>(in, out, tab, c->dev);
        launched(c, "k_mulplain_g");
        CU_TRY(cudaGetLastError());
        return FHE_OK;
    }
    if (int rc = set_smem(k_mulplain<LOGN, LOGE>, smem)) return rc;
    k_mulplain<LOGN, LOGE><<<(unsigned)npoly, S::T, smem, st>>>(in, out, tab, c->dev);
    launched(c, "k_mulplain");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}

// mode: NTT_OUT_CANON ([0, q), the public transform and the plaintext tables) or
// NTT_OUT_REDUCED ([0, 4q), input range of the pointwise DCT kernel)
static int ntt_forward(fhe_ctx *c, const u64 *in, u64 *out, uint64_t npoly, int mode, cudaStream_t st)
{
    if (!npoly) return FHE_OK;
    if (npoly > 0x7fffffffull) return fail(FHE_ERR_INVALID_ARGUMENT, "too many polynomials in one call");
#define CALL(LN, LE) return launch_fwd<LN, LE>(c, in, out, npoly, mode, st)
    FHE_DISPATCH_LOGN(c->host.logn, CALL)
#undef CALL
    return FHE_OK;
}
// scale: multiply by n^-1 (public transform) or not (the preceding table already carries n^-1)
static int ntt_inverse(fhe_ctx *c, const u64 *in, u64 *out, uint64_t npoly, bool scale, cudaStream_t st)
{
    if (!npoly) return FHE_OK;
    if (npoly > 0x7fffffffull) return fail(FHE_ERR_INVALID_ARGUMENT, "too many polynomials in one call");
#define CALL(LN, LE)                                                        \
    if (scale) return launch_inv<LN, LE, true>(c, in, out, npoly, st);      \
    return launch_inv<LN, LE, false>(c, in, out, npoly, st)
    FHE_DISPATCH_LOGN(c->host.logn, CALL)
#undef CALL
    return FHE_OK;
}

static int check_ctx(const fhe_ctx *c)
{
    if (!c) return fail(FHE_ERR_INVALID_ARGUMENT, "null context");
    CU_TRY(cudaSetDevice(c->device));
    return FHE_OK;
}

extern "C" {

const char *fhe_last_error(void) { return g_err.c_str(); }
int fhe_version(void) { return 100; }

int fhe_default_coeff_modulus(uint32_t n, uint64_t *q_out, uint32_t *k_out)
{
    // SEAL 2.3.1 coeff_modulus_128 ([SEAL-recall]): the largest primes below 2^55
    // (2^54 for n = 2048) that are 1 mod 2^18, descending.
    static const uint64_t P[8] = {0x7fffffff380001ull, 0x7ffffffef00001ull, 0x7ffffffeac0001ull,
                                  0x7ffffffe700001ull, 0x7ffffffe600001ull, 0x7ffffffe4c0001ull,
                                  0x7ffffffdec0001ull, 0x7ffffffdd00001ull};
    if (!q_out || !k_out) return fail(FHE_ERR_INVALID_ARGUMENT, "null output");
    uint32_t k = 0;
    switch (n) {
    case 2048: q_out[0] = 0x3fffffff000001ull; *k_out = 1; return FHE_OK;
    case 4096: k = 2; break;
    case 8192: k = 4; break;
    case 16384: k = 8; break;
    default: return fail(FHE_ERR_INVALID_ARGUMENT, "coeff_modulus_128 is defined for n in {2048, 4096, 8192, 16384}");
    }
    for (uint32_t i = 0; i < k; ++i) q_out[i] = P[i];
    *k_out = k;
    return FHE_OK;
}

int fhe_ctx_create(uint32_t n, uint32_t k, const uint64_t *q, uint64_t t, int device, void *cuda_stream,
                   fhe_ctx **out)
{
    if (!out || !q) return fail(FHE_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    auto c = std::make_unique<fhe_ctx>();
    std::string err;
    if (!build_ring(c->host, n, k, q, t, err)) return fail(FHE_ERR_INVALID_ARGUMENT, err);
    c->generic = c->host.generic;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(FHE_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(FHE_ERR_INVALID_ARGUMENT, "device index out of range");
    c->device = device;
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(FHE_ERR_CUDA, "libfhe_b200 is built for sm_100a only");
    c->sm_count = prop.multiProcessorCount;
    c->alloc_cache_limit = prop.totalGlobalMem / 2;  // blocks parked for reuse: at most half the device
    {
        // keep freed blocks in the stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cuda_stream) c->stream = (cudaStream_t)cuda_stream;
    else {
        CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    CU_TRY(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < fhe_ctx::kSlots; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
    CU_TRY(cudaEventCreateWithFlags(&c->ev_mark, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
    for (int i = 0; i < fhe_ctx::kDown; ++i) CU_TRY(cudaEventCreateWithFlags(&c->ev_down[i], cudaEventDisableTiming));
    for (int i = 0; i < fhe_ctx::kDown; ++i) CU_TRY(cudaEventCreateWithFlags(&c->ev_upr[i], cudaEventDisableTiming));
    const size_t tbytes = (size_t)k * n * sizeof(W2);
    CU_TRY(cudaMalloc(&c->d_tw, tbytes));
    CU_TRY(cudaMalloc(&c->d_itw, tbytes));
    CU_TRY(cudaMalloc(&c->d_scratch, 64));
    std::vector<W2> flat((size_t)k * n);
    for (uint32_t j = 0; j < k; ++j) std::copy(c->host.limbs[j].tw.begin(), c->host.limbs[j].tw.end(), flat.begin() + (size_t)j * n);
    CU_TRY(cudaMemcpy(c->d_tw, flat.data(), tbytes, cudaMemcpyHostToDevice));
    for (uint32_t j = 0; j < k; ++j) std::copy(c->host.limbs[j].itw.begin(), c->host.limbs[j].itw.end(), flat.begin() + (size_t)j * n);
    CU_TRY(cudaMemcpy(c->d_itw, flat.data(), tbytes, cudaMemcpyHostToDevice));
    RingDev &D = c->dev;
    D.n = n; D.logn = c->host.logn; D.k = k; D.t = t; D.plain_thr = c->host.plain_thr; D.q_mod_t = c->host.q_mod_t;
    for (uint32_t j = 0; j < k; ++j) {
        D.limb[j] = c->host.limbs[j].dev;
        D.delta[j] = c->host.limbs[j].delta;
        D.ratio0[j] = c->host.limbs[j].ratio0;
        D.ratio1[j] = c->host.limbs[j].ratio1;
    }
    D.tw = c->d_tw; D.itw = c->d_itw;
    *out = c.release();
    return FHE_OK;
}

void fhe_ctx_destroy(fhe_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < fhe_ctx::kSlots; ++i) {
        if (c->stage[i]) cudaFree(c->stage[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    clear_stage_events(c);
    if (c->ev_mark) cudaEventDestroy(c->ev_mark);
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    for (int i = 0; i < fhe_ctx::kDown; ++i) {
        if (c->ev_down[i]) cudaEventDestroy(c->ev_down[i]);
        if (c->ev_upr[i]) cudaEventDestroy(c->ev_upr[i]);
    }
    for (auto &m : c->prof_marks) cudaEventDestroy(m.second);
    for (auto &kv : c->alloc_size) cudaFreeAsync(kv.first, c->stream);  // parked and still-live blocks alike
    cudaStreamSynchronize(c->stream);
    behz_state_destroy(c->behz);
    cudaFree(c->d_tw); cudaFree(c->d_itw); cudaFree(c->d_scratch);
    if (c->h_ring) {
        cudaFreeHost(c->h_ring);
        for (int i = 0; i < fhe_ctx::kRing; ++i) cudaEventDestroy(c->ev_ring[i]);
    }
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

int fhe_ctx_sync(fhe_ctx *c)
{
    if (int rc = check_ctx(c)) return rc;
    CU_TRY(cudaStreamSynchronize(c->stream));
    return FHE_OK;
}

size_t fhe_ct_bytes(const fhe_ctx *c, uint32_t size) { return c ? (size_t)size * c->host.k * c->host.n * sizeof(u64) : 0; }
uint64_t fhe_launch_count(const fhe_ctx *c) { return c ? c->launches : 0; }
int fhe_set_dct_chunk_blocks(fhe_ctx *c, uint32_t blocks)
{
    if (!c) return fail(FHE_ERR_INVALID_ARGUMENT, "null context");
    c->dct_chunk_blocks = blocks;
    return FHE_OK;
}

int fhe_set_stage_timing(fhe_ctx *c, int enabled)
{
    if (!c) return fail(FHE_ERR_INVALID_ARGUMENT, "null context");
    c->stage_timing = enabled != 0;
    if (!enabled) clear_stage_events(c);
    return FHE_OK;
}
int fhe_get_stage_times(fhe_ctx *c, float *ms)
{
    if (int rc = check_ctx(c)) return rc;
    if (!ms) return fail(FHE_ERR_INVALID_ARGUMENT, "null output");
    CU_TRY(cudaStreamSynchronize(c->stream));
    ms[0] = ms[1] = ms[2] = 0.f;
    for (size_t i = 0; i + 3 < c->stage_ev.size(); i += 4)
        for (int s = 0; s < 3; ++s) {
            float t = 0.f;
            CU_TRY(cudaEventElapsedTime(&t, c->stage_ev[i + s], c->stage_ev[i + s + 1]));
            ms[s] += t;
        }
    return FHE_OK;
}

int fhe_profile_begin(fhe_ctx *c)
{
    if (int rc = check_ctx(c)) return rc;
    for (auto &m : c->prof_marks) cudaEventDestroy(m.second);
    c->prof_marks.clear();
    c->prof_on = true;
    launched(c, "");  // the start mark (not a kernel)
    c->launches--;
    return FHE_OK;
}
int fhe_profile_end(fhe_ctx *c, char *names, size_t names_len, float *ms, uint64_t *counts, uint32_t max_kernels, uint32_t *n_out)
{
    if (int rc = check_ctx(c)) return rc;
    if (!names || !ms || !counts || !n_out) return fail(FHE_ERR_INVALID_ARGUMENT, "null output");
    c->prof_on = false;
    CU_TRY(cudaStreamSynchronize(c->stream));
    std::vector<std::string> order;
    std::unordered_map<std::string, std::pair<float, uint64_t>> acc;
    for (size_t i = 1; i < c->prof_marks.size(); ++i) {
        float t = 0.f;
        CU_TRY(cudaEventElapsedTime(&t, c->prof_marks[i - 1].second, c->prof_marks[i].second));
        const std::string k = c->prof_marks[i].first;
        if (!acc.count(k)) order.push_back(k);
        acc[k].first += t;
        acc[k].second += 1;
    }
    for (auto &m : c->prof_marks) cudaEventDestroy(m.second);
    c->prof_marks.clear();
    std::string joined;
    uint32_t n = 0;
    for (const std::string &k : order) {
        if (n == max_kernels) break;
        if (joined.size() + k.size() + 2 > names_len) break;
        joined += k;
        joined += '\n';
        ms[n] = acc[k].first;
        counts[n] = acc[k].second;
        ++n;
    }
    std::memcpy(names, joined.c_str(), joined.size() + 1);
    *n_out = n;
    return FHE_OK;
}

int fhe_device_alloc(fhe_ctx *c, size_t bytes, void **dptr)
{
    if (int rc = check_ctx(c)) return rc;
    if (!dptr) return fail(FHE_ERR_INVALID_ARGUMENT, "null output");
    // The seal-shaped host layer creates and drops a temporary ciphertext (or a batch of them, up
    // to gigabytes) per Evaluator call.  cudaMalloc / cudaFree would synchronise the device every
    // time; the driver's stream-ordered pool does not, but growing it in multi-gigabyte steps stalls
    // for hundreds of milliseconds at unpredictable moments (measured: 0.24 s and 1.4 s for the same
    // decode pass).  So freed blocks are kept in exact-size free lists and handed out again without
    // touching the driver: safe because every buffer of a context lives on its one stream.
    const size_t b = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
    auto it = c->alloc_free.find(b);
    if (it != c->alloc_free.end() && !it->second.empty()) {
        *dptr = it->second.back();
        it->second.pop_back();
        c->alloc_cached -= b;
        c->alloc_size[*dptr] = b;  // clears the "parked" mark
        return FHE_OK;
    }
    cudaError_t e = cudaMallocAsync(dptr, b, c->stream);
    if (e != cudaSuccess) {
        // out of memory with blocks parked in the cache: give them back and try once more
        cudaGetLastError();
        for (auto &kv : c->alloc_free) {
            for (void *p : kv.second) { cudaFreeAsync(p, c->stream); c->alloc_size.erase(p); }
            kv.second.clear();
        }
        c->alloc_cached = 0;
        cudaStreamSynchronize(c->stream);
        e = cudaMallocAsync(dptr, b, c->stream);
    }
    if (e != cudaSuccess) return fail(FHE_ERR_ALLOC, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    c->alloc_size[*dptr] = b;
    return FHE_OK;
}
int fhe_device_free(fhe_ctx *c, void *dptr)
{
    if (int rc = check_ctx(c)) return rc;
    if (!dptr) return FHE_OK;
    auto it = c->alloc_size.find(dptr);
    if (it == c->alloc_size.end()) return fail(FHE_ERR_INVALID_ARGUMENT, "fhe_device_free: not a pointer of fhe_device_alloc on this context");
    if (it->second & 1) return fail(FHE_ERR_INVALID_ARGUMENT, "fhe_device_free: block freed twice");
    const size_t b = it->second;  // multiple of 256: bit 0 marks a parked block
    if (c->alloc_cached + b <= c->alloc_cache_limit) {
        c->alloc_free[b].push_back(dptr);
        c->alloc_cached += b;
        it->second = b | 1;
        return FHE_OK;
    }
    c->alloc_size.erase(it);
    CU_TRY(cudaFreeAsync(dptr, c->stream));
    return FHE_OK;
}
int fhe_host_alloc(size_t bytes, void **hptr)
{
    if (!hptr) return fail(FHE_ERR_INVALID_ARGUMENT, "null output");
    cudaError_t e = cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(FHE_ERR_ALLOC, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    return FHE_OK;
}
int fhe_host_free(void *hptr)
{
    CU_TRY(cudaFreeHost(hptr));
    return FHE_OK;
}
int fhe_upload(fhe_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (int rc = check_ctx(c)) return rc;
    CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return FHE_OK;
}
int fhe_download(fhe_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (int rc = check_ctx(c)) return rc;
    CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return FHE_OK;
}

// Asynchronous transfers on the context's two copy streams.  A transfer starts after everything queued on
// the context's stream at the time of the call (the destination of an upload may be a recycled block that
// earlier kernels still read; the source of a download is produced by them) and runs beside whatever is
// queued afterwards.  fhe_transfer_fence: the context's STREAM waits for all uploads issued so far (call
// it before the first kernel that reads an uploaded buffer); the HOST waits until all downloads but the
// `keep_downloads` most recent ones have landed (call it before freeing or overwriting the source of an
// older download: the host blocks, the device keeps working on what is already queued).
static int wait_downloads(fhe_ctx *c, uint64_t keep)
{
    while (c->down_waited + keep < c->down_issued) {
        CU_TRY(cudaEventSynchronize(c->ev_down[c->down_waited % fhe_ctx::kDown]));
        ++c->down_waited;
    }
    return FHE_OK;
}
static int wait_uploads(fhe_ctx *c, uint64_t keep)
{
    while (c->up_waited + keep < c->up_issued) {
        CU_TRY(cudaEventSynchronize(c->ev_upr[c->up_waited % fhe_ctx::kDown]));
        ++c->up_waited;
    }
    return FHE_OK;
}
int fhe_upload_fence(fhe_ctx *c, uint32_t keep_uploads)
{
    if (int rc = check_ctx(c)) return rc;
    return wait_uploads(c, keep_uploads);
}
int fhe_upload_async(fhe_ctx *c, void *d_dst, const void *h_src, size_t bytes)
{
    if (int rc = check_ctx(c)) return rc;
    if (!d_dst || !h_src) return fail(FHE_ERR_INVALID_ARGUMENT, "null buffer");
    if (int rc = wait_uploads(c, fhe_ctx::kDown - 1)) return rc;  // the event about to be reused has completed
    CU_TRY(cudaEventRecord(c->ev_mark, c->stream));
    CU_TRY(cudaStreamWaitEvent(c->s_in, c->ev_mark, 0));
    CU_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->s_in));
    CU_TRY(cudaEventRecord(c->ev_up, c->s_in));
    c->up_pending = true;
    CU_TRY(cudaEventRecord(c->ev_upr[c->up_issued % fhe_ctx::kDown], c->s_in));
    ++c->up_issued;
    return FHE_OK;
}
int fhe_download_async(fhe_ctx *c, void *h_dst, const void *d_src, size_t bytes)
{
    if (int rc = check_ctx(c)) return rc;
    if (!h_dst || !d_src) return fail(FHE_ERR_INVALID_ARGUMENT, "null buffer");
    if (int rc = wait_downloads(c, fhe_ctx::kDown - 1)) return rc;  // the event about to be reused has completed
    CU_TRY(cudaEventRecord(c->ev_mark, c->stream));
    CU_TRY(cudaStreamWaitEvent(c->s_out, c->ev_mark, 0));
    CU_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->s_out));
    CU_TRY(cudaEventRecord(c->ev_down[c->down_issued % fhe_ctx::kDown], c->s_out));
    ++c->down_issued;
    return FHE_OK;
}
int fhe_transfer_fence(fhe_ctx *c, uint32_t keep_downloads)
{
    if (int rc = check_ctx(c)) return rc;
    if (c->up_pending) CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_up, 0));
    return wait_downloads(c, keep_downloads);
}

int fhe_device_copy(fhe_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (int rc = check_ctx(c)) return rc;
    if (!dst || !src) return fail(FHE_ERR_INVALID_ARGUMENT, "null buffer");
    CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return FHE_OK;
}
int fhe_device_copy_2d(fhe_ctx *c, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows)
{
    if (int rc = check_ctx(c)) return rc;
    if (!dst || !src || width > dst_pitch || width > src_pitch) return fail(FHE_ERR_INVALID_ARGUMENT, "bad 2-D copy");
    if (!rows || !width) return FHE_OK;
    CU_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, rows, cudaMemcpyDeviceToDevice, c->stream));
    return FHE_OK;
}

// FractionalEncoder(t, poly_modulus, int_coeffs, frac_coeffs, 2).encode:
// integer part in binary from x^0 up; i-th fractional binary digit d_i (weight 2^-(i+1)) stored
// as -d_i at x^(n-1-i) because x^n = -1; negative values negate every digit; all mod t.
int fhe_frac_encode(uint32_t n, uint64_t t_, double value, uint32_t int_coeffs, uint32_t frac_coeffs,
                    uint32_t base, uint64_t *plain_out)
{
    if (!plain_out) return fail(FHE_ERR_INVALID_ARGUMENT, "null argument");
    if (base != 2) return fail(FHE_ERR_UNSUPPORTED, "FractionalEncoder base must be 2");
    if (n < 2 || (n & (n - 1)) || t_ < 2) return fail(FHE_ERR_INVALID_ARGUMENT, "bad ring degree or plain modulus");
    const u64 t = t_;
    if ((uint64_t)int_coeffs + frac_coeffs > n) return fail(FHE_ERR_INVALID_ARGUMENT, "integer + fraction coefficient counts exceed n");
    if (!(value == value) || value > 9.2e18 || value < -9.2e18) return fail(FHE_ERR_INVALID_ARGUMENT, "value out of range");
    std::memset(plain_out, 0, sizeof(uint64_t) * n);
    const bool neg = value < 0;
    const u64 plus = neg ? t - 1 : 1, minus = neg ? 1 : t - 1;
    double mag = neg ? -value : value;
    u64 ip = (u64)mag;  // truncation
    double frac = mag - (double)ip;
    for (uint32_t i = 0; ip; ++i, ip >>= 1) {
        if (i >= int_coeffs) return fail(FHE_ERR_INVALID_ARGUMENT, "integer part does not fit int_coeffs");
        if (ip & 1) plain_out[i] = plus;
    }
    for (uint32_t i = 0; i < frac_coeffs && frac != 0; ++i) {
        frac *= 2.0;
        if (frac >= 1.0) {
            frac -= 1.0;
            plain_out[n - 1 - i] = minus;
        }
    }
    return FHE_OK;
}

static int elementwise(fhe_ctx *c, int op, u64 *a, const u64 *b, uint64_t n_ct, uint32_t size)
{
    if (int rc = check_ctx(c)) return rc;
    if (!a || (op != 2 && !b) || size == 0) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    const unsigned long long total = (unsigned long long)n_ct * size * c->host.k * c->host.n;
    if (!total) return FHE_OK;
    const int g = grid_for(total, 256, c->sm_count);
    if (op == 0) k_elementwise<0><<<g, 256, 0, c->stream>>>(a, b, total, c->dev);
    else if (op == 1) k_elementwise<1><<<g, 256, 0, c->stream>>>(a, b, total, c->dev);
    else k_elementwise<2><<<g, 256, 0, c->stream>>>(a, nullptr, total, c->dev);
    launched(c, "k_elementwise");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}
int fhe_add(fhe_ctx *c, uint64_t *a, const uint64_t *b, uint64_t n_ct, uint32_t size) { return elementwise(c, 0, (u64 *)a, (const u64 *)b, n_ct, size); }
int fhe_sub(fhe_ctx *c, uint64_t *a, const uint64_t *b, uint64_t n_ct, uint32_t size) { return elementwise(c, 1, (u64 *)a, (const u64 *)b, n_ct, size); }
int fhe_negate(fhe_ctx *c, uint64_t *a, uint64_t n_ct, uint32_t size) { return elementwise(c, 2, (u64 *)a, nullptr, n_ct, size); }

uint32_t fhe_ntt_loge(uint32_t n)
{
    switch (n) {
    case 1024: case 2048: case 4096: return 4;
    case 8192: return FHE_LOGE13;
    case 16384: return FHE_LOGE14;
    default: return 0;
    }
}
int fhe_ntt_forward(fhe_ctx *c, uint64_t *a, uint64_t npoly)
{
    if (int rc = check_ctx(c)) return rc;
    if (!a) return fail(FHE_ERR_INVALID_ARGUMENT, "null buffer");
    return ntt_forward(c, (const u64 *)a, (u64 *)a, npoly, NTT_OUT_CANON, c->stream);
}
int fhe_ntt_inverse(fhe_ctx *c, uint64_t *a, uint64_t npoly)
{
    if (int rc = check_ctx(c)) return rc;
    if (!a) return fail(FHE_ERR_INVALID_ARGUMENT, "null buffer");
    return ntt_inverse(c, (const u64 *)a, (u64 *)a, npoly, true, c->stream);
}

// plaintexts (host [count][n]) -> device W2 table [count][k][n] of NTT(lift(plain)) (optionally
// times a second table, period b_period) and, if d_scaled, the Delta-scaled polys.
static int build_plain_tables(fhe_ctx *c, const uint64_t *plain, uint32_t count, W2 *d_ntt, int ntt_times_ninv,
                              u64 *d_scaled, u64 *d_values_out /* optional [count][k][n] reduced NTT values */)
{
    const uint32_t n = c->host.n, k = c->host.k;
    for (size_t i = 0; i < (size_t)count * n; ++i)
        if (plain[i] >= c->host.t) return fail(FHE_ERR_INVALID_ARGUMENT, "plaintext coefficient is not below plain_modulus");
    u64 *d_plain = nullptr, *d_lift = nullptr;
    const unsigned long long total = (unsigned long long)count * k * n;
    CU_TRY(cudaMalloc(&d_plain, (size_t)count * n * sizeof(u64)));
    cudaError_t e = cudaMalloc(&d_lift, total * sizeof(u64));
    if (e != cudaSuccess) { cudaFree(d_plain); return fail(FHE_ERR_ALLOC, "cudaMalloc failed for plaintext tables"); }
    int rc = FHE_OK;
    do {
        if (cudaMemcpyAsync(d_plain, plain, (size_t)count * n * sizeof(u64), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = fail(FHE_ERR_CUDA, "plaintext upload failed"); break; }
        const int g = grid_for(total, 256, c->sm_count);
        if (d_scaled) { k_plain_scale<<<g, 256, 0, c->stream>>>(d_plain, d_scaled, total, c->dev); launched(c, "k_plain_scale"); }
        k_plain_lift<<<g, 256, 0, c->stream>>>(d_plain, d_lift, total, c->dev);
        launched(c, "k_plain_lift");
        if ((rc = ntt_forward(c, d_lift, d_lift, (uint64_t)count * k, NTT_OUT_CANON, c->stream))) break;
        if (d_ntt) { k_make_shoup<<<g, 256, 0, c->stream>>>(d_lift, nullptr, d_ntt, total, 1, ntt_times_ninv, c->dev, c->generic ? 64 : 63); launched(c, "k_make_shoup"); }
        if (d_values_out && cudaMemcpyAsync(d_values_out, d_lift, total * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) { rc = fail(FHE_ERR_CUDA, "table copy failed"); break; }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = fail(FHE_ERR_CUDA, "plaintext table kernels failed"); break; }
    } while (0);
    cudaFree(d_plain);
    cudaFree(d_lift);
    return rc;
}

// dst[pl][0][j][c] = Delta-scaled plaintext (as k_plain_scale), dst[pl][1] = 0
__global__ void k_encrypt_transparent(const u64 *plain, u64 *dst, unsigned long long total, RingDev R)
{
    // total = count * 2 * k * n; i = ((pl * 2 + poly) * k + j) * n + c
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i & (R.n - 1));
        const unsigned long long lp = i >> R.logn;            // (pl * 2 + poly) * k + j
        const uint32_t j = (uint32_t)(lp % R.k);
        const unsigned long long pp = lp / R.k;               // pl * 2 + poly
        if (pp & 1) { dst[i] = 0; continue; }
        const u64 m = plain[(pp >> 1) * R.n + c];
        u64 lo = R.delta[j] * m, hi = mulhi_u64(R.delta[j], m);
        if (m >= R.plain_thr) {
            u64 s = lo + R.q_mod_t;
            hi += (s < lo);
            lo = s;
        }
        dst[i] = barrett128(lo, hi, R.limb[j].a.q, R.ratio0[j], R.ratio1[j]);
    }
}

// Small host -> device parameter blocks (pointer tables of the resize graphs, the plaintexts of
// server-side encryptions, short per-element plaintexts) WITHOUT the copy engine: the block is written into
// a slot range of mapped pinned memory and the kernel reads it from there over PCIe.  A cudaMemcpyAsync
// would neither synchronise nor cost much by itself -- but every host-to-device copy of the process shares
// one copy-engine queue, so a 24 KB table queued behind an 800 MB row upload of the streaming resize
// pipeline (fhe_upload_async) held its kernel back for the whole upload (measured: end-to-end resize 2.1 k
// pixels/s with the uploads against 3.9 k without, profiles/r02_notes.md F).
// stage_params: copies `bytes` into free slots (waits on the host only if the ring has wrapped onto slots
// whose consumer has not finished) and returns the device address; stage_release: call after the LAST kernel
// that reads the block has been launched.
struct StageRef {
    void *dptr = nullptr;
    int first = 0, count = 0;
};
static int stage_params(fhe_ctx *c, const void *src, size_t bytes, StageRef *ref)
{
    const size_t slot = fhe_ctx::kStageSlot;
    if (!c->h_ring) {
        void *h = nullptr, *d = nullptr;
        CU_TRY(cudaHostAlloc(&h, (size_t)fhe_ctx::kRing * slot, cudaHostAllocMapped | cudaHostAllocPortable));
        CU_TRY(cudaHostGetDevicePointer(&d, h, 0));
        c->h_ring = (char *)h;
        c->d_ring = (char *)d;
        for (int i = 0; i < fhe_ctx::kRing; ++i) CU_TRY(cudaEventCreateWithFlags(&c->ev_ring[i], cudaEventDisableTiming));
    }
    const int need = (int)((bytes + slot - 1) / slot);
    if (need < 1 || need > fhe_ctx::kRing / 2) return fail(FHE_ERR_INVALID_ARGUMENT, "parameter block too large for the staging ring");
    if (c->ring_next + need > fhe_ctx::kRing) c->ring_next = 0;
    ref->first = c->ring_next;
    ref->count = need;
    for (int i = ref->first; i < ref->first + need; ++i)
        if (c->ring_used[i]) CU_TRY(cudaEventSynchronize(c->ev_ring[i]));
    memcpy(c->h_ring + (size_t)ref->first * slot, src, bytes);
    ref->dptr = c->d_ring + (size_t)ref->first * slot;
    c->ring_next = ref->first + need;
    return FHE_OK;
}
static int stage_release(fhe_ctx *c, const StageRef &ref)
{
    for (int i = ref.first; i < ref.first + ref.count; ++i) {
        CU_TRY(cudaEventRecord(c->ev_ring[i], c->stream));
        c->ring_used[i] = true;
    }
    return FHE_OK;
}

int fhe_encrypt_transparent(fhe_ctx *c, const uint64_t *plain, uint32_t count, uint64_t *d_dst)
{
    if (int rc = check_ctx(c)) return rc;
    if (!plain || !d_dst || !count) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    const uint32_t n = c->host.n, k = c->host.k;
    for (size_t i = 0; i < (size_t)count * n; ++i)
        if (plain[i] >= c->host.t) return fail(FHE_ERR_INVALID_ARGUMENT, "plaintext coefficient is not below plain_modulus");
    // the plaintexts are staged (no stream synchronisation, no copy-engine traffic: see stage_params), a
    // group at a time, and the kernel reads them in place; `plain` may be reused by the caller at once
    const size_t pbytes = (size_t)n * sizeof(u64);
    const uint32_t group = (uint32_t)std::max<size_t>(1, ((size_t)fhe_ctx::kRing / 4 * fhe_ctx::kStageSlot) / pbytes);
    for (uint32_t i = 0; i < count; i += group) {
        const uint32_t m = count - i < group ? count - i : group;
        StageRef ref;
        if (int rc = stage_params(c, plain + (size_t)i * n, m * pbytes, &ref)) return rc;
        const unsigned long long total = (unsigned long long)m * 2 * k * n;
        k_encrypt_transparent<<<grid_for(total, 256, c->sm_count), 256, 0, c->stream>>>((const u64 *)ref.dptr, (u64 *)d_dst + (size_t)i * 2 * k * n, total, c->dev);
        launched(c, "k_encrypt_transparent");
        CU_TRY(cudaGetLastError());
        if (int rc = stage_release(c, ref)) return rc;
    }
    return FHE_OK;
}

int fhe_ptab_create(fhe_ctx *c, const uint64_t *plain, uint32_t count, fhe_ptab **out)
{
    if (int rc = check_ctx(c)) return rc;
    if (!plain || !out || !count) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    *out = nullptr;
    auto t = std::make_unique<fhe_ptab>();
    t->ctx = c;
    t->count = count;
    const uint32_t n = c->host.n, k = c->host.k;
    t->is_zero.assign(count, 1);
    for (uint32_t p = 0; p < count; ++p)
        for (uint32_t i = 0; i < n; ++i)
            if (plain[(size_t)p * n + i]) { t->is_zero[p] = 0; break; }
    t->sparse.assign(count, {});
    for (uint32_t p = 0; p < count; ++p) {
        std::vector<uint32_t> terms;
        bool ok = true;
        for (uint32_t i = 0; i < n && ok; ++i) {
            const uint64_t m = plain[(size_t)p * n + i];
            if (!m) continue;
            // t - 1 is the lift of -1 (for t = 2 it coincides with 1 and is negative: m >= plain_thr)
            if (m == c->host.t - 1) terms.push_back(i | 0x80000000u);
            else if (m == 1) terms.push_back(i);
            else ok = false;
            if (terms.size() > (size_t)SPARSE_MAX) ok = false;
        }
        if (ok && !terms.empty()) t->sparse[p] = std::move(terms);
    }
    const size_t total = (size_t)count * k * n;
    CU_TRY(cudaMalloc(&t->d_ntt, total * sizeof(W2)));
    cudaError_t e = cudaMalloc(&t->d_scaled, total * sizeof(u64));
    if (e != cudaSuccess) { cudaFree(t->d_ntt); return fail(FHE_ERR_ALLOC, "cudaMalloc failed"); }
    if (int rc = build_plain_tables(c, plain, count, t->d_ntt, 1, t->d_scaled, nullptr)) {
        cudaFree(t->d_ntt); cudaFree(t->d_scaled);
        return rc;
    }
    *out = t.release();
    return FHE_OK;
}
void fhe_ptab_destroy(fhe_ptab *t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaFree(t->d_ntt);
    cudaFree(t->d_scaled);
    delete t;
}

int fhe_multiply_plain(fhe_ctx *c, uint64_t *a, uint64_t n_ct, uint32_t size, const fhe_ptab *tab, uint32_t index)
{
    if (int rc = check_ctx(c)) return rc;
    if (!a || !tab || tab->ctx != c || index >= tab->count || !size) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    if (tab->is_zero[index]) return fail(FHE_ERR_ZERO_PLAIN, "plain cannot be zero");
    const uint64_t npoly = n_ct * size * c->host.k;
    if (!npoly) return FHE_OK;
    if (npoly > 0x7fffffffull) return fail(FHE_ERR_INVALID_ARGUMENT, "too many polynomials in one call");
    if (!tab->sparse[index].empty()) {
        SparseTerms T{};
        T.n_terms = (uint32_t)tab->sparse[index].size();
        for (uint32_t i = 0; i < T.n_terms; ++i) T.term[i] = tab->sparse[index][i];
        const size_t smem = (size_t)c->host.n * sizeof(u64);
        if (int rc = set_smem(k_mulplain_sparse, smem)) return rc;
        k_mulplain_sparse<<<(unsigned)npoly, 256, smem, c->stream>>>((u64 *)a, T, c->dev);
        launched(c, "k_mulplain_sparse");
        CU_TRY(cudaGetLastError());
        return FHE_OK;
    }
    const W2 *table = tab->d_ntt + (size_t)index * c->host.k * c->host.n;
#define CALL(LN, LE) return launch_mulplain<LN, LE>(c, (const u64 *)a, (u64 *)a, table, npoly, c->stream)
    FHE_DISPATCH_LOGN(c->host.logn, CALL)
#undef CALL
    return FHE_OK;
}

static int plain_addsub(fhe_ctx *c, int sign, uint64_t *a, uint64_t n_ct, uint32_t size, const fhe_ptab *tab, uint32_t index)
{
    if (int rc = check_ctx(c)) return rc;
    if (!a || !tab || tab->ctx != c || index >= tab->count || !size) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    const unsigned long long total = (unsigned long long)n_ct * c->host.k * c->host.n;
    if (!total) return FHE_OK;
    const u64 *scaled = tab->d_scaled + (size_t)index * c->host.k * c->host.n;
    const int g = grid_for(total, 256, c->sm_count);
    if (sign > 0) k_plain_addsub<1><<<g, 256, 0, c->stream>>>((u64 *)a, scaled, n_ct, size, c->dev);
    else k_plain_addsub<-1><<<g, 256, 0, c->stream>>>((u64 *)a, scaled, n_ct, size, c->dev);
    launched(c, "k_plain_addsub");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}
int fhe_add_plain(fhe_ctx *c, uint64_t *a, uint64_t n_ct, uint32_t size, const fhe_ptab *tab, uint32_t index) { return plain_addsub(c, +1, a, n_ct, size, tab, index); }
int fhe_sub_plain(fhe_ctx *c, uint64_t *a, uint64_t n_ct, uint32_t size, const fhe_ptab *tab, uint32_t index) { return plain_addsub(c, -1, a, n_ct, size, tab, index); }

int fhe_dct_plan_create(fhe_ctx *c, const uint64_t *dct_consts, const uint64_t *eighth, const uint64_t *quant, fhe_dct_plan **out)
{
    if (int rc = check_ctx(c)) return rc;
    if (!dct_consts || !eighth || !out) return fail(FHE_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (c->generic)
        return fail(FHE_ERR_UNSUPPORTED, "the fused DCT needs every coeff_modulus prime in [2^40, 2^55) (SEAL 2.3's coeff_modulus_128 defaults are); "
                                         "with this chain evaluate encrypted_dct through the Evaluator entry points");
    const uint32_t n = c->host.n, k = c->host.k;
    // Evaluator::multiply_plain throws on a zero plaintext: so does plan creation.
    auto zero = [&](const uint64_t *p) { for (uint32_t i = 0; i < n; ++i) if (p[i]) return false; return true; };
    for (int i = 0; i < DCT_NCONST; ++i) if (zero(dct_consts + (size_t)i * n)) return fail(FHE_ERR_ZERO_PLAIN, "plain cannot be zero");
    if (zero(eighth)) return fail(FHE_ERR_ZERO_PLAIN, "plain cannot be zero");
    if (quant) for (int i = 0; i < 64; ++i) if (zero(quant + (size_t)i * n)) return fail(FHE_ERR_ZERO_PLAIN, "plain cannot be zero");
    auto p = std::make_unique<fhe_dct_plan>();
    p->ctx = c;
    p->has_quant = quant != nullptr;
    const size_t per = (size_t)k * n;
    u64 *d_e = nullptr, *d_q = nullptr;
    CU_TRY(cudaMalloc(&p->d_c, DCT_NCONST * per * sizeof(W2)));
    CU_TRY(cudaMalloc(&p->d_s, 64 * per * sizeof(W2)));
    CU_TRY(cudaMalloc(&d_e, per * sizeof(u64)));
    CU_TRY(cudaMalloc(&d_q, 64 * per * sizeof(u64)));
    int rc = build_plain_tables(c, dct_consts, DCT_NCONST, p->d_c, 0, nullptr, nullptr);
    if (!rc) rc = build_plain_tables(c, eighth, 1, nullptr, 0, nullptr, d_e);
    if (!rc && quant) rc = build_plain_tables(c, quant, 64, nullptr, 0, nullptr, d_q);
    if (!rc) {
        // merged scale constants: s[i] = NTT(0.125) * NTT(1/quant[i]) (ring multiplication is
        // associative, so (x * 0.125) * (1/quant) == x * (0.125 * 1/quant) exactly), or NTT(0.125);
        // times n^-1, which the inverse NTT after the pointwise kernel then does not apply again.
        const unsigned long long total = 64ull * per;
        const int g = grid_for(total, 256, c->sm_count);
        if (quant) k_make_shoup<<<g, 256, 0, c->stream>>>(d_q, d_e, p->d_s, total, per, 1, c->dev);
        else {
            // replicate the eighth table 64 times
            for (int i = 0; i < 64; ++i) cudaMemcpyAsync(d_q + (size_t)i * per, d_e, per * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream);
            k_make_shoup<<<g, 256, 0, c->stream>>>(d_q, nullptr, p->d_s, total, 1, 1, c->dev);
        }
        launched(c, "k_make_shoup");
        if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = fail(FHE_ERR_CUDA, "scale table kernel failed");
    }
    cudaFree(d_e);
    cudaFree(d_q);
    if (rc) { cudaFree(p->d_c); cudaFree(p->d_s); return rc; }
    *out = p.release();
    return FHE_OK;
}
void fhe_dct_plan_destroy(fhe_dct_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaFree(p->d_c);
    cudaFree(p->d_s);
    delete p;
}

// Pointwise DCT (+quantize) of `nblocks` blocks whose first ciphertexts are `blk_stride`
// ciphertexts apart (64 = dense; 192 = one channel of RGB tiles), in place on `data`, an array of
// `n_ct` size-2 ciphertexts.
static int launch_dct_pointwise(fhe_ctx *c, const fhe_dct_plan *plan, u64 *data, uint64_t nblocks, uint32_t blk_stride,
                                uint64_t n_ct, cudaStream_t st)
{
    const uint32_t n = c->host.n, k = c->host.k;
    const unsigned long long nunits = nblocks * 2;
    // units per CTA: 32 amortise the constant loads of a CTA; fewer when that would leave the
    // grid under a few waves of 2 CTAs per SM
    const unsigned long long per_ug = (unsigned long long)k * (n / DCT_TILE), want = 8ull * c->sm_count;
    int upc = 32;
    while (upc > 1 && ((nunits + upc - 1) / upc) * per_ug < want) upc >>= 1;
    const unsigned long long grid = ((nunits + upc - 1) / upc) * per_ug;
    if (grid > 0x7fffffffull) return fail(FHE_ERR_INVALID_ARGUMENT, "too many blocks in one chunk");
    CUtensorMap tm;
    if (int rc = make_ct_tensor_map(&tm, data, n_ct, k, n)) return rc;
    if (int rc = set_smem(k_dct_pointwise, DCT_SMEM)) return rc;
    k_dct_pointwise<<<(unsigned)grid, 256, DCT_SMEM, st>>>(tm, plan->d_c, plan->d_s, c->dev, nunits, upc, blk_stride);
    launched(c, "k_dct_pointwise");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}

// three kernels per chunk: forward NTT (in -> out), pointwise DCT in place, inverse NTT in place
static int dct_chunk(fhe_ctx *c, const fhe_dct_plan *plan, const u64 *in, u64 *out, uint64_t nblocks, cudaStream_t st)
{
    const uint32_t k = c->host.k;
    const uint64_t npoly = nblocks * 64 * 2 * k;
    auto mark = [&]() -> int {
        if (!c->stage_timing) return FHE_OK;
        cudaEvent_t e;
        CU_TRY(cudaEventCreate(&e));
        c->stage_ev.push_back(e);
        CU_TRY(cudaEventRecord(e, st));
        return FHE_OK;
    };
    if (int rc = mark()) return rc;
    if (int rc = ntt_forward(c, in, out, npoly, NTT_OUT_REDUCED, st)) return rc;
    if (int rc = mark()) return rc;
    if (int rc = launch_dct_pointwise(c, plan, out, nblocks, 64, nblocks * 64, st)) return rc;
    if (int rc = mark()) return rc;
    if (int rc = ntt_inverse(c, out, out, npoly, false, st)) return rc;
    return mark();
}

static const uint64_t kMaxChunkBlocks = 1ull << 20;

int fhe_dct8x8_quantize(fhe_ctx *c, const fhe_dct_plan *plan, const uint64_t *d_in, uint64_t *d_out, uint64_t nblocks)
{
    if (int rc = check_ctx(c)) return rc;
    if (!plan || plan->ctx != c || !d_in || !d_out) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    const size_t blk_words = (size_t)64 * 2 * c->host.k * c->host.n;
    clear_stage_events(c);
    uint64_t chunk = c->dct_chunk_blocks ? c->dct_chunk_blocks : nblocks;
    if (chunk > kMaxChunkBlocks / c->host.k) chunk = kMaxChunkBlocks / c->host.k;
    for (uint64_t b = 0; b < nblocks; b += chunk) {
        const uint64_t nb = nblocks - b < chunk ? nblocks - b : chunk;
        if (int rc = dct_chunk(c, plan, (const u64 *)d_in + b * blk_words, (u64 *)d_out + b * blk_words, nb, c->stream)) return rc;
    }
    return FHE_OK;
}

int fhe_dct8x8_quantize_host(fhe_ctx *c, const fhe_dct_plan *plan, const uint64_t *h_in, uint64_t *h_out, uint64_t nblocks)
{
    if (int rc = check_ctx(c)) return rc;
    if (!plan || plan->ctx != c || !h_in || !h_out) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    if (!nblocks) return FHE_OK;
    clear_stage_events(c);
    const size_t blk_bytes = (size_t)64 * 2 * c->host.k * c->host.n * sizeof(u64);
    // staging slots of up to 64 MiB: the copies bound this path (PCIe), so small slots keep the
    // pipeline's fill and drain (one H2D and one D2H that overlap nothing) short
    uint64_t sb = (64ull << 20) / blk_bytes;
    if (sb < 1) sb = 1;
    if (sb > nblocks) sb = nblocks;
    if (c->stage_blocks < sb) {
        for (int i = 0; i < fhe_ctx::kSlots; ++i) {
            if (c->stage[i]) cudaFree(c->stage[i]);
            c->stage[i] = nullptr;
            cudaError_t e = cudaMalloc(&c->stage[i], sb * blk_bytes);
            if (e != cudaSuccess) { c->stage_blocks = 0; return fail(FHE_ERR_ALLOC, "cudaMalloc failed for staging"); }
        }
        c->stage_blocks = sb;
    }
    CU_TRY(cudaStreamSynchronize(c->stream));
    uint64_t idx = 0;
    for (uint64_t b = 0; b < nblocks; b += sb, ++idx) {
        const int s = (int)(idx % fhe_ctx::kSlots);
        const uint64_t nb = nblocks - b < sb ? nblocks - b : sb;
        const char *src = (const char *)h_in + b * blk_bytes;
        char *dst = (char *)h_out + b * blk_bytes;
        if (idx >= (uint64_t)fhe_ctx::kSlots) CU_TRY(cudaStreamWaitEvent(c->s_in, c->ev_out[s], 0));  // slot drained
        CU_TRY(cudaMemcpyAsync(c->stage[s], src, nb * blk_bytes, cudaMemcpyHostToDevice, c->s_in));
        CU_TRY(cudaEventRecord(c->ev_in[s], c->s_in));
        CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[s], 0));
        if (int rc = dct_chunk(c, plan, c->stage[s], c->stage[s], nb, c->stream)) return rc;
        CU_TRY(cudaEventRecord(c->ev_comp[s], c->stream));
        CU_TRY(cudaStreamWaitEvent(c->s_out, c->ev_comp[s], 0));
        CU_TRY(cudaMemcpyAsync(dst, c->stage[s], nb * blk_bytes, cudaMemcpyDeviceToHost, c->s_out));
        CU_TRY(cudaEventRecord(c->ev_out[s], c->s_out));
    }
    CU_TRY(cudaStreamSynchronize(c->s_out));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return FHE_OK;
}

int fhe_fill_uniform(fhe_ctx *c, uint64_t *d_dst, uint64_t first_ct, uint64_t n_ct, uint32_t size, uint64_t seed)
{
    if (int rc = check_ctx(c)) return rc;
    if (!d_dst || !size) return fail(FHE_ERR_INVALID_ARGUMENT, "bad argument");
    const unsigned long long total = (unsigned long long)n_ct * size * c->host.k * c->host.n;
    if (!total) return FHE_OK;
    k_fill_uniform<<<grid_for(total, 256, c->sm_count), 256, 0, c->stream>>>((u64 *)d_dst, first_ct, total, size, seed, c->dev);
    launched(c, "k_fill_uniform");
    CU_TRY(cudaGetLastError());
    return FHE_OK;
}

int fhe_checksum(fhe_ctx *c, const uint64_t *d, uint64_t count, uint64_t *out)
{
    if (int rc = check_ctx(c)) return rc;
    if (!d || !out) return fail(FHE_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaMemsetAsync(c->d_scratch, 0, 8, c->stream));
    if (count) {
        k_checksum<<<grid_for(count, 256, c->sm_count), 256, 0, c->stream>>>((const u64 *)d, count, c->d_scratch);
        launched(c, "k_checksum");
    }
    unsigned long long h = 0;
    CU_TRY(cudaMemcpyAsync(&h, c->d_scratch, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    *out = h;
    return FHE_OK;
}

}  // extern "C"

#include "behz_api.inc"
#include "resize_api.inc"
#include "jpeg_api.inc"
