// emu.cpp -- CPU emulation of the CUDA thread bodies (tests only).
//
// There is no GPU in the development container, so the __host__ __device__ kernel bodies in
// csrc/ntt.cuh and csrc/dct.cuh are executed here thread by thread (all tau of a pass, then
// the next pass = what __syncthreads() orders on the device).  This checks the index math,
// twiddle indexing, lazy-range bookkeeping and table conventions against the oracle before any
// GPU time is spent.  It is NOT a product code path: nothing in the package loads this library.
#include <cstring>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "../../fully-homomorphic-image-processing_b200/csrc/behz_host.hpp"
#include "../../fully-homomorphic-image-processing_b200/csrc/dct.cuh"
#include "../../fully-homomorphic-image-processing_b200/csrc/ntt.cuh"
#include "../../fully-homomorphic-image-processing_b200/csrc/ring_host.hpp"

using namespace fhe;

template <int V> using ic = std::integral_constant<int, V>;

template <int LOGN, int LOGE>
static void emu_fwd(const HostLimb &L, const u64 *in, u64 *out, int mode)
{
    using S = NttShape<LOGN, LOGE>;
    std::vector<u64> smem(S::SMEM_WORDS), tmp(in, in + S::N);
    auto pass = [&](auto tag) {
        constexpr int P = decltype(tag)::value;
        for (int tau = 0; tau < S::T; ++tau) {
            if (mode == NTT_OUT_CANON) ntt_fwd_pass<LOGN, LOGE, P, NTT_OUT_CANON>(tau, tmp.data(), out, smem.data(), L.tw.data(), L.dev);
            else if (mode == NTT_OUT_REDUCED) ntt_fwd_pass<LOGN, LOGE, P, NTT_OUT_REDUCED>(tau, tmp.data(), out, smem.data(), L.tw.data(), L.dev);
            else ntt_fwd_pass<LOGN, LOGE, P, NTT_OUT_LAZY>(tau, tmp.data(), out, smem.data(), L.tw.data(), L.dev);
        }
    };
    pass(ic<0>{});
    if constexpr (S::NPASS > 1) pass(ic<1>{});
    if constexpr (S::NPASS > 2) pass(ic<2>{});
    if constexpr (S::NPASS > 3) pass(ic<3>{});
}
template <int LOGN, int LOGE>
static void emu_inv(const HostLimb &L, const u64 *in, u64 *out, bool scale)
{
    using S = NttShape<LOGN, LOGE>;
    std::vector<u64> smem(S::SMEM_WORDS), tmp(in, in + S::N);
    auto pass = [&](auto tag) {
        constexpr int P = decltype(tag)::value;
        for (int tau = 0; tau < S::T; ++tau) {
            if (scale) ntt_inv_pass<LOGN, LOGE, P, true>(tau, tmp.data(), out, smem.data(), L.itw.data(), L.dev);
            else ntt_inv_pass<LOGN, LOGE, P, false>(tau, tmp.data(), out, smem.data(), L.itw.data(), L.dev);
        }
    };
    if constexpr (S::NPASS > 3) pass(ic<3>{});
    if constexpr (S::NPASS > 2) pass(ic<2>{});
    if constexpr (S::NPASS > 1) pass(ic<1>{});
    pass(ic<0>{});
}

static int fwd_limb(const HostLimb &L, uint32_t logn, const u64 *in, u64 *out, int mode)
{
    switch (logn) {
    case 10: emu_fwd<10, 4>(L, in, out, mode); return 0;
    case 11: emu_fwd<11, 4>(L, in, out, mode); return 0;
    case 12: emu_fwd<12, 4>(L, in, out, mode); return 0;
    case 13: emu_fwd<13, FHE_LOGE13>(L, in, out, mode); return 0;
    case 14: emu_fwd<14, FHE_LOGE14>(L, in, out, mode); return 0;
    default: return -1;
    }
}
static int fwd_dispatch(const HostRing &R, int j, const u64 *in, u64 *out, int mode) { return fwd_limb(R.limbs[j], R.logn, in, out, mode); }
static int inv_limb(const HostLimb &L, uint32_t logn, const u64 *in, u64 *out, bool scale)
{
    switch (logn) {
    case 10: emu_inv<10, 4>(L, in, out, scale); return 0;
    case 11: emu_inv<11, 4>(L, in, out, scale); return 0;
    case 12: emu_inv<12, 4>(L, in, out, scale); return 0;
    case 13: emu_inv<13, FHE_LOGE13>(L, in, out, scale); return 0;
    case 14: emu_inv<14, FHE_LOGE14>(L, in, out, scale); return 0;
    default: return -1;
    }
}

static int inv_dispatch(const HostRing &R, int j, const u64 *in, u64 *out, bool scale) { return inv_limb(R.limbs[j], R.logn, in, out, scale); }
// the B_sk limbs run the same transforms as the q limbs (csrc/behz.cuh: 55-bit auxiliary primes)
static int fwd_w_dispatch(const HostBehz &H, int j, const u64 *in, u64 *out) { return fwd_limb(H.limbs[j], H.dev.logn, in, out, NTT_OUT_CANON); }
static int inv_w_dispatch(const HostBehz &H, int j, const u64 *in, u64 *out) { return inv_limb(H.limbs[j], H.dev.logn, in, out, true); }

extern "C" {

// ---- BEHZ multiply path (csrc/behz.cuh bodies) ----
void *emu_behz_create(void *ring, char *err, int errlen)
{
    auto H = std::make_unique<HostBehz>();
    std::string e;
    if (!build_behz(*(HostRing *)ring, *H, e)) {
        if (err && errlen > 0) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
        return nullptr;
    }
    return H.release();
}
void emu_behz_destroy(void *h) { delete (HostBehz *)h; }
uint32_t emu_behz_nb(void *h) { return ((HostBehz *)h)->dev.nb; }
uint64_t emu_behz_bsk(void *h, int j) { return ((HostBehz *)h)->dev.b[j]; }
int emu_nttw_forward(void *h, int j, const uint64_t *in, uint64_t *out) { return fwd_w_dispatch(*(HostBehz *)h, j, (const u64 *)in, (u64 *)out); }
int emu_nttw_inverse(void *h, int j, const uint64_t *in, uint64_t *out) { return inv_w_dispatch(*(HostBehz *)h, j, (const u64 *)in, (u64 *)out); }
// x [k][n] -> out [nb][n]
void emu_behz_extend(void *h, const uint64_t *x, uint64_t *out)
{
    const BehzDev &B = ((HostBehz *)h)->dev;
    for (uint32_t c = 0; c < B.n; ++c) behz_extend_coeff(B, (const u64 *)x + c, B.n, (u64 *)out + c, B.n);
}
void emu_behz_scale(void *h, const uint64_t *xq, const uint64_t *xb, uint64_t *out)
{
    const BehzDev &B = ((HostBehz *)h)->dev;
    for (uint32_t c = 0; c < B.n; ++c) behz_scale_coeff(B, (const u64 *)xq + c, B.n, (const u64 *)xb + c, B.n, (u64 *)out + c, B.n);
}
// the whole multiply as the library sequences it: extend, forward NTTs, tensor, inverse NTTs, scale
int emu_multiply(void *ring, void *h, const uint64_t *a1, uint32_t s1, const uint64_t *a2, uint32_t s2, uint64_t *dest)
{
    const HostRing &R = *(HostRing *)ring;
    const HostBehz &H = *(HostBehz *)h;
    const BehzDev &B = H.dev;
    const uint32_t n = B.n, k = B.k, nb = B.nb, sd = s1 + s2 - 1;
    const size_t pq = (size_t)k * n, pb = (size_t)nb * n;
    std::vector<u64> q1(s1 * pq), b1(s1 * pb), q2(s2 * pq), b2(s2 * pb), dq(sd * pq), db(sd * pb);
    auto prep = [&](const uint64_t *src, uint32_t s, std::vector<u64> &tq, std::vector<u64> &tb) {
        for (uint32_t p = 0; p < s; ++p) {
            emu_behz_extend(h, src + p * pq, (uint64_t *)tb.data() + p * pb);
            for (uint32_t i = 0; i < k; ++i) fwd_dispatch(R, i, (const u64 *)src + p * pq + (size_t)i * n, tq.data() + p * pq + (size_t)i * n, NTT_OUT_CANON);
            for (uint32_t j = 0; j < nb; ++j) fwd_w_dispatch(H, j, tb.data() + p * pb + (size_t)j * n, tb.data() + p * pb + (size_t)j * n);
        }
    };
    prep(a1, s1, q1, b1);
    prep(a2, s2, q2, b2);
    for (uint32_t i = 0; i < k; ++i)
        for (uint32_t c = 0; c < n; ++c)
            behz_tensor_elem(q1.data() + (size_t)i * n + c, q2.data() + (size_t)i * n + c, dq.data() + (size_t)i * n + c, pq, s1, s2, B.q[i], B.qr0[i], B.qr1[i], B.qdelta[i]);
    for (uint32_t j = 0; j < nb; ++j)
        for (uint32_t c = 0; c < n; ++c)
            behz_tensor_elem(b1.data() + (size_t)j * n + c, b2.data() + (size_t)j * n + c, db.data() + (size_t)j * n + c, pb, s1, s2, B.b[j], B.br0[j], B.br1[j], B.bdelta[j]);
    for (uint32_t p = 0; p < sd; ++p) {
        for (uint32_t i = 0; i < k; ++i) inv_dispatch(R, i, dq.data() + p * pq + (size_t)i * n, dq.data() + p * pq + (size_t)i * n, true);
        for (uint32_t j = 0; j < nb; ++j) inv_w_dispatch(H, j, db.data() + p * pb + (size_t)j * n, db.data() + p * pb + (size_t)j * n);
        emu_behz_scale(h, (const uint64_t *)dq.data() + p * pq, (const uint64_t *)db.data() + p * pb, dest + p * pq);
    }
    return 0;
}

// z = z1 2^64 + z0 (< 2^119) mod p (p = 2^55 - delta), the special-prime reduction of csrc/behz.cuh
uint64_t emu_fold_aux(uint64_t z0, uint64_t z1, uint64_t p)
{
    return fold_aux(((u128d)z1 << 64) | z0, p, (u32)((1ull << 55) - p));
}

void *emu_ring_create(uint32_t n, uint32_t k, const uint64_t *q, uint64_t t, char *err, int errlen)
{
    auto R = std::make_unique<HostRing>();
    std::string e;
    if (!build_ring(*R, n, k, q, t, e) || (R->generic && !(e = "generic chain: the emulation covers the lazy kernels only").empty())) {
        if (err && errlen > 0) { std::strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
        return nullptr;
    }
    return R.release();
}
void emu_ring_destroy(void *r) { delete (HostRing *)r; }
uint64_t emu_psi(void *r, int j) { return ((HostRing *)r)->limbs[j].psi; }
uint64_t emu_delta(void *r, int j) { return ((HostRing *)r)->limbs[j].delta; }
uint64_t emu_q_mod_t(void *r) { return ((HostRing *)r)->q_mod_t; }

// mode: NTT_OUT_LAZY (0), NTT_OUT_REDUCED (1: [0, 4q)), NTT_OUT_CANON (2: [0, q))
int emu_ntt_forward(void *r, int j, const uint64_t *in, uint64_t *out, int mode)
{
    return fwd_dispatch(*(HostRing *)r, j, (const u64 *)in, (u64 *)out, mode);
}
// input values < 5q; scale != 0 multiplies by n^-1
int emu_ntt_inverse(void *r, int j, const uint64_t *in, uint64_t *out, int scale)
{
    return inv_dispatch(*(HostRing *)r, j, (const u64 *)in, (u64 *)out, scale != 0);
}
// x*w lazily (mulc), reduce63 and canon of single values, for range tests
uint64_t emu_mulc(void *r, int j, uint64_t x, uint64_t w)
{
    const HostLimb &L = ((HostRing *)r)->limbs[j];
    return mulc(x, h_shoup(w, L.q), L.dev.a.n2q);
}
uint64_t emu_reduce63(void *r, int j, uint64_t x) { return reduce63(x, ((HostRing *)r)->limbs[j].dev.a); }
uint64_t emu_canon(void *r, int j, uint64_t x) { return canon(x, ((HostRing *)r)->limbs[j].dev.a); }

// Emulated fused DCT(+quantize) on `nblocks` blocks: forward NTT of every limb-poly, pointwise
// row/column phases per (unit, limb, tile), inverse NTT.  plains: [12][n] DCT constants,
// eighth: [n], quant: [64][n] or NULL.  data: [block][64][2][k][n], in place.
int emu_dct_quant(void *r, uint64_t *data_, uint64_t nblocks, const uint64_t *plains,
                  const uint64_t *eighth, const uint64_t *quant)
{
    const HostRing &R = *(HostRing *)r;
    u64 *data = (u64 *)data_;
    const uint32_t n = R.n, k = R.k;
    // plan tables in the permuted NTT domain
    auto lift_ntt = [&](const uint64_t *plain, uint32_t j, std::vector<u64> &out) {
        out.assign(n, 0);
        for (uint32_t i = 0; i < n; ++i) out[i] = plain[i] >= R.plain_thr ? plain[i] + (R.limbs[j].q - R.t) : plain[i];
        fwd_dispatch(R, j, out.data(), out.data(), NTT_OUT_CANON);
    };
    std::vector<std::vector<W2>> ctab(k), stab(k);
    std::vector<u64> a, b;
    for (uint32_t j = 0; j < k; ++j) {
        u64 q = R.limbs[j].q;
        ctab[j].resize((size_t)DCT_NCONST * n);
        stab[j].resize((size_t)64 * n);
        for (int c = 0; c < DCT_NCONST; ++c) {
            lift_ntt(plains + (size_t)c * n, j, a);
            for (uint32_t i = 0; i < n; ++i) ctab[j][(size_t)c * n + i] = h_shoup(a[i], q);
        }
        lift_ntt(eighth, j, a);
        for (int s = 0; s < 64; ++s) {
            if (quant) lift_ntt(quant + (size_t)s * n, j, b);
            for (uint32_t i = 0; i < n; ++i)
                stab[j][(size_t)s * n + i] = h_shoup(h_mulmod(quant ? h_mulmod(a[i], b[i], q) : a[i], R.limbs[j].dev.ninv.w, q), q);
        }
    }
    const size_t ct_stride = (size_t)2 * k * n;
    for (uint64_t blk = 0; blk < nblocks; ++blk) {
        u64 *B = data + blk * 64 * ct_stride;
        for (size_t lp = 0; lp < (size_t)64 * 2 * k; ++lp)
            fwd_dispatch(R, lp % k, B + lp * n, B + lp * n, NTT_OUT_REDUCED);
        std::vector<u64> sbuf(64 * DCT_TILE);
        for (uint32_t p = 0; p < 2; ++p)
            for (uint32_t j = 0; j < k; ++j)
                for (uint32_t tile = 0; tile < n / DCT_TILE; ++tile) {
                    u64 *base = B + ((size_t)p * k + j) * n + (size_t)tile * DCT_TILE;
                    u64 q = R.limbs[j].q, n2q = R.limbs[j].dev.a.n2q;
                    for (int w = 0; w < 8; ++w)
                        for (int lane = 0; lane < DCT_TILE; ++lane) {
                            W2 c[DCT_NCONST];
                            for (int ci = 0; ci < DCT_NCONST; ++ci) c[ci] = ctab[j][(size_t)ci * n + tile * DCT_TILE + lane];
                            dct_row_phase(w, lane, base, ct_stride, sbuf.data(), c, q, n2q);
                        }
                    for (int w = 0; w < 8; ++w)
                        for (int lane = 0; lane < DCT_TILE; ++lane) {
                            W2 c[DCT_NCONST], s[8];
                            for (int ci = 0; ci < DCT_NCONST; ++ci) c[ci] = ctab[j][(size_t)ci * n + tile * DCT_TILE + lane];
                            for (int i = 0; i < 8; ++i) s[i] = stab[j][(size_t)(8 * i + w) * n + tile * DCT_TILE + lane];
                            dct_col_phase(w, lane, base, ct_stride, sbuf.data(), c, s, 1, q, n2q);
                        }
                }
        for (size_t lp = 0; lp < (size_t)64 * 2 * k; ++lp)
            inv_dispatch(R, lp % k, B + lp * n, B + lp * n, false);
    }
    return 0;
}
}
