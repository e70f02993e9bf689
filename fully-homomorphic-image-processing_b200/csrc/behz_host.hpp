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

// The auxiliary primes: 2^55 - j 2^15 + 1 for j = 1, 2, 3 ... (1 mod 2^15, so every supported ring
// degree has its 2n-th roots), prime, j not a multiple of 8 (SEAL's own coeff_modulus primes are 1 mod
// 2^18: the defaults can never collide), and different from every prime of q.  The first is m_sk,
// the next kb form B.  (SEAL 2.3 uses its 61-bit internal primes here; the result does not depend
// on the base -- see behz.cuh.)
inline std::vector<u64> h_aux_primes(size_t count, const std::vector<u64> &q)
{
    std::vector<u64> out;
    for (u64 j = 1; out.size() < count && j < 2048; ++j) {
        if (j % 8 == 0) continue;
        const u64 p = (1ull << 55) - (j << 15) + 1;
        bool clash = false;
        for (u64 qi : q) clash |= qi == p;
        if (!clash && h_is_prime(p)) out.push_back(p);
    }
    return out;
}

struct HostBehz {
    BehzDev dev{};
    std::vector<HostLimb> limbs;  // B_sk, m_sk last: the same tables and constants as a limb of q
};

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
    // B m_sk must exceed 2 |W|, |W| < t n 32 q (behz.cuh); SEAL's rule 32 + bits(t) + bits(q) < bits(B_sk) is
    // kept (it is the stricter one) with this base's 55-bit primes
    if (32 + tbits + total_bits >= 55 * k + 55) ++kb;
    if (kb + 1 > (uint32_t)BEHZ_MAXB) { err = "auxiliary base too large"; return false; }
    const uint32_t nb = kb + 1;
    const std::vector<u64> aux = h_aux_primes(nb, q);
    if (aux.size() < nb) { err = "could not find the auxiliary primes"; return false; }
    const u64 kMsk = aux[0];
    std::vector<u64> B(kb), bsk(nb);
    for (uint32_t i = 0; i < kb; ++i) B[i] = bsk[i] = aux[1 + i];
    bsk[kb] = kMsk;
    BehzDev &D = H.dev;
    D = BehzDev{};
    D.n = n; D.logn = R.logn; D.k = k; D.kb = kb; D.nb = nb;
    const u64 MT = 1ull << 32;
    const size_t none = (size_t)-1;
    for (uint32_t i = 0; i < k; ++i) {
        D.q[i] = q[i];
        D.qr0[i] = R.limbs[i].ratio0; D.qr1[i] = R.limbs[i].ratio1;
        D.qdelta[i] = fold_delta_of(q[i]);
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
    H.limbs.assign(nb, HostLimb{});
    for (uint32_t j = 0; j < nb; ++j) {
        const u64 p = bsk[j];
        D.b[j] = p;
        const u128 ratio = (~(u128)0) / p;
        D.br0[j] = (u64)ratio; D.br1[j] = (u64)(ratio >> 64);
        D.bdelta[j] = (u32)((1ull << 55) - p);
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
        if (!h_build_limb(H.limbs[j], p, n, R.logn, err)) return false;
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
