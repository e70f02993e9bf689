// behz_host.hpp -- host-side precomputation of the BEHZ constants (product code, runs once per
// context): the counterpart of SEAL 2.3 util/baseconverter BaseConverter::generate ([SEAL-recall],
// SURVEY.md 0.1 / Appendix A.6).  Auxiliary base: |B| = k primes (k + 1 when
// 32 + bits(t) + bits(q) >= 61 k + 61), B_sk = B u {m_sk}, m~ = 2^32.
#pragma once
#include <string>
#include <vector>

#include "behz.cuh"
#include "ring_host.hpp"

namespace fhe {

// SEAL 2.3.1 internal_mods: the largest primes below 2^61 that are 1 mod 2^18, descending
// (m_sk, gamma -- decryption only, unused here --, then aux_small_mods).
static const u64 kMsk = 0x1fffffffffe00001ull;
static const u64 kAuxSmallMods[10] = {0x1fffffffffb40001ull, 0x1fffffffff500001ull, 0x1fffffffff380001ull,
                                      0x1fffffffff000001ull, 0x1ffffffffef00001ull, 0x1ffffffffee80001ull,
                                      0x1ffffffffeb40001ull, 0x1ffffffffe780001ull, 0x1ffffffffe600001ull,
                                      0x1ffffffffe4c0001ull};

struct HostBehz {
    BehzDev dev{};
    std::vector<LimbWide> limbs;            // B_sk
    std::vector<std::vector<W2>> tw, itw;   // per B_sk prime, [n] per-pass layout, Shoup pairs scaled by 2^64
};

inline W2 h_shoup64(u64 w, u64 q) { return W2{w, (u64)((((u128)w) << 64) / q)}; }
inline u64 h_prod_except(const std::vector<u64> &list, size_t skip, u64 p)
{
    u64 r = 1 % p;
    for (size_t i = 0; i < list.size(); ++i)
        if (i != skip) r = h_mulmod(r, list[i] % p, p);
    return r;
}

inline bool build_behz(const HostRing &R, HostBehz &H, std::string &err)
{
    const uint32_t n = R.n, k = R.k;
    if (k > (uint32_t)BEHZ_MAXK) { err = "Evaluator::multiply supports up to 8 coeff_modulus primes in this build"; return false; }
    std::vector<u64> q(k);
    uint32_t total_bits = 0;
    for (uint32_t i = 0; i < k; ++i) {
        q[i] = R.limbs[i].q;
        total_bits += 64 - __builtin_clzll(q[i]);
    }
    const uint32_t tbits = 64 - __builtin_clzll(R.t);
    uint32_t kb = k;
    if (32 + tbits + total_bits >= 61 * k + 61) ++kb;
    if (kb + 1 > (uint32_t)BEHZ_MAXB) { err = "auxiliary base too large"; return false; }
    const uint32_t nb = kb + 1;
    std::vector<u64> B(kb), bsk(nb);
    for (uint32_t i = 0; i < kb; ++i) B[i] = bsk[i] = kAuxSmallMods[i];
    bsk[kb] = kMsk;
    for (u64 p : bsk) {
        if ((p - 1) % (2ull * n)) { err = "auxiliary prime is not 1 mod 2n"; return false; }
        for (u64 qi : q)
            if (qi == p) { err = "coeff_modulus collides with the auxiliary base"; return false; }
    }
    BehzDev &D = H.dev;
    D = BehzDev{};
    D.n = n; D.logn = R.logn; D.k = k; D.kb = kb; D.nb = nb;
    const u64 MT = 1ull << 32;
    const size_t none = (size_t)-1;
    for (uint32_t i = 0; i < k; ++i) {
        D.q[i] = q[i];
        D.qr0[i] = R.limbs[i].ratio0; D.qr1[i] = R.limbs[i].ratio1;
        const u64 qhat_inv = h_invmod(h_prod_except(q, i, q[i]), q[i]);
        D.mt_qhat_inv[i] = h_shoup64(h_mulmod(qhat_inv, MT % q[i], q[i]), q[i]);
        D.qhat_mod_mt[i] = (u32)h_prod_except(q, i, MT);
        D.T[i] = h_shoup64(h_mulmod(qhat_inv, R.t % q[i], q[i]), q[i]);
        const u64 Bq = h_prod_except(B, none, q[i]);
        D.Bq_pos[i] = Bq;
        D.Bq_neg[i] = Bq ? q[i] - Bq : 0;
        for (uint32_t l = 0; l < kb; ++l) D.I[i][l] = h_prod_except(B, l, q[i]);
    }
    {
        u64 qm = h_prod_except(q, none, MT), x = 1;  // Newton: x = q^-1 mod 2^32
        for (int it = 0; it < 6; ++it) x = (x * (2 - qm * x)) & (MT - 1);
        D.neg_inv_q_mod_mt = (u32)(0 - x);
    }
    H.limbs.assign(nb, LimbWide{});
    H.tw.assign(nb, {}); H.itw.assign(nb, {});
    for (uint32_t j = 0; j < nb; ++j) {
        const u64 p = bsk[j];
        D.b[j] = p;
        const u128 ratio = (~(u128)0) / p;
        D.br0[j] = (u64)ratio; D.br1[j] = (u64)(ratio >> 64);
        if (p >= (1ull << 61) || (1ull << 61) - p >= (1ull << 26)) { err = "auxiliary prime is not 2^61 - delta with delta < 2^26"; return false; }
        D.bd8[j] = (u32)(8 * ((1ull << 61) - p));
        const u64 inv_mt = h_invmod(MT % p, p);
        const u64 q_mod = h_prod_except(q, none, p), q_inv = h_invmod(q_mod, p);
        const u64 dpos = h_mulmod(q_mod, inv_mt, p);
        D.Dpos[j] = dpos; D.Dneg[j] = dpos ? p - dpos : 0;
        D.E[j] = h_mulmod(R.t % p, q_inv, p);
        for (uint32_t i = 0; i < k; ++i) {
            const u64 qhat = h_prod_except(q, i, p);
            D.C[j][i] = h_mulmod(qhat, inv_mt, p);
            const u64 f = h_mulmod(qhat, q_inv, p);
            D.F[j][i] = f ? p - f : 0;
        }
        // NTT tables of this prime
        LimbWide &L = H.limbs[j];
        L.q = p; L.q2 = 2 * p; L.r0 = D.br0[j]; L.r1 = D.br1[j];
        const u64 psi = h_min_primitive_root(p, n);
        if (!psi) { err = "no primitive 2n-th root for an auxiliary prime"; return false; }
        std::vector<u64> pw(n);
        pw[0] = 1;
        for (uint32_t i = 1; i < n; ++i) pw[i] = h_mulmod(pw[i - 1], psi, p);
        std::vector<W2> seal_tw(n), seal_itw(n);
        for (uint32_t idx = 0; idx < n; ++idx) {
            const u64 w = pw[h_bitrev(idx, R.logn)];
            seal_tw[idx] = h_shoup64(w, p);
            seal_itw[idx] = h_shoup64(h_invmod(w, p), p);
        }
        H.tw[j].assign(n, W2{0, 0}); H.itw[j].assign(n, W2{0, 0});
        switch (R.logn) {
        case 10: ntt_permute_table<10, 4>(seal_tw.data(), H.tw[j].data()); ntt_permute_table<10, 4>(seal_itw.data(), H.itw[j].data()); break;
        case 11: ntt_permute_table<11, 4>(seal_tw.data(), H.tw[j].data()); ntt_permute_table<11, 4>(seal_itw.data(), H.itw[j].data()); break;
        case 12: ntt_permute_table<12, 4>(seal_tw.data(), H.tw[j].data()); ntt_permute_table<12, 4>(seal_itw.data(), H.itw[j].data()); break;
        case 13: ntt_permute_table<13, FHE_LOGE13>(seal_tw.data(), H.tw[j].data()); ntt_permute_table<13, FHE_LOGE13>(seal_itw.data(), H.itw[j].data()); break;
        case 14: ntt_permute_table<14, FHE_LOGE14>(seal_tw.data(), H.tw[j].data()); ntt_permute_table<14, FHE_LOGE14>(seal_itw.data(), H.itw[j].data()); break;
        }
        const u64 ninv = h_invmod(n % p, p);
        L.ninv = h_shoup64(ninv, p);
        L.ninv_w1 = h_shoup64(h_mulmod(ninv, seal_itw[1].w, p), p);
    }
    const u64 invB = h_invmod(h_prod_except(B, none, kMsk), kMsk);
    D.H = kMsk - invB;
    for (uint32_t l = 0; l < kb; ++l) {
        D.bhat_inv[l] = h_shoup64(h_invmod(h_prod_except(B, l, B[l]), B[l]), B[l]);
        D.G[l] = h_mulmod(h_prod_except(B, l, kMsk), invB, kMsk);
    }
    return true;
}

}  // namespace fhe
