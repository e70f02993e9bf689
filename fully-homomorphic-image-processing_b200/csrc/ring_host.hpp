// ring_host.hpp -- host-side precomputation for an encryption context (n, {q_j}, t).
//
// Product code (not the oracle): the B200-side counterpart of what seal::SEALContext derives
// from EncryptionParameters at server_jpeg.cpp:74-80 -- per-limb NTT twiddles (psi = minimal
// primitive 2n-th root, powers in bit-reversed order, Shoup companions), Barrett ratios,
// Delta = floor(q/t) mod q_j, q mod t, n^-1.  Runs once per context on the CPU (parameter
// setup, not the hot path); the tables are then uploaded and never touched by the host again.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ntt.cuh"

namespace fhe {

typedef unsigned __int128 u128;

struct HostLimb {
    u64 q = 0, psi = 0, ratio0 = 0, ratio1 = 0, delta = 0;
    std::vector<W2> tw, itw;  // [n], per-pass layout (ntt.cuh PassGeom::tw_index); n-1 entries used
    u64 itw1 = 0;             // psi^-(n/2): SEAL-order inverse table entry 1
    LimbDev dev{};
};

struct HostRing {
    uint32_t n = 0, logn = 0, k = 0;
    u64 t = 0, q_mod_t = 0, plain_thr = 0;
    int max_qbits = 0;
    bool generic = false;     // a prime outside [2^40, 2^55): every limb runs the Harvey arithmetic of ntt_generic.cuh
    std::vector<HostLimb> limbs;
};

inline u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
inline u64 h_powmod(u64 b, u64 e, u64 q)
{
    u64 r = 1;
    b %= q;
    for (; e; e >>= 1, b = h_mulmod(b, b, q))
        if (e & 1) r = h_mulmod(r, b, q);
    return r;
}
inline u64 h_invmod(u64 a, u64 q) { return h_powmod(a, q - 2, q); }
// multiplier constant with its mulc companion floor(w 2^63 / q)  (modarith.cuh)
inline W2 h_shoup(u64 w, u64 q) { return W2{w, (u64)((((u128)w) << 63) / q)}; }
// ... and with the exact Shoup companion floor(w 2^64 / q) (ntt_generic.cuh mulsh, behz.cuh)
inline W2 h_shoup64(u64 w, u64 q) { return W2{w, (u64)((((u128)w) << 64) / q)}; }
inline LimbArith h_limb_arith(u64 q)
{
    LimbArith a;
    a.q = q; a.n2q = 0 - 2 * q; a.nq = 0 - q; a.q5 = 5 * q;
    for (int i = 0; i < 4; ++i) { a.kq5[i] = (5 * q) << i; a.kq16[i] = (16 * q) << i; }
    a.mu63 = (u32)((((u128)1) << 63) / q);
    a.mu64 = (u32)((((u128)1) << 64) / q);
    return a;
}

inline bool h_is_prime(u64 n)
{
    if (n < 2) return false;
    for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (n % p == 0) return n == p;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1)) d >>= 1, ++r;
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        u64 x = h_powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r && comp; ++i) {
            x = h_mulmod(x, x, n);
            if (x == n - 1) comp = false;
        }
        if (comp) return false;
    }
    return true;
}

inline uint32_t h_bitrev(uint32_t x, uint32_t bits)
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// smallest primitive 2n-th root of unity mod q (the convention SEAL 2.3 uses [SEAL-recall];
// unobservable for every op that returns to coefficient form).
inline u64 h_min_primitive_root(u64 q, uint32_t n)
{
    u64 root = 0;
    for (u64 g = 2; g < 4096 && !root; ++g) {
        u64 r = h_powmod(g, (q - 1) / (2ull * n), q);
        if (h_powmod(r, n, q) == q - 1) root = r;
    }
    if (!root) return 0;
    u64 sq = h_mulmod(root, root, q), cur = root, best = root;
    for (uint32_t i = 0; i < n; ++i, cur = h_mulmod(cur, sq, q))
        if (cur < best) best = cur;
    return best;
}

// NTT tables and arithmetic constants of one prime (1 mod 2n); generic: Shoup companions scaled by 2^64
// for ntt_generic.cuh instead of mulc's 2^63 (which needs q < 2^55)
inline bool h_build_limb(HostLimb &L, u64 qj, uint32_t n, uint32_t logn, std::string &err, bool generic = false)
{
    auto shoup = [generic](u64 w, u64 q) { return generic ? h_shoup64(w, q) : h_shoup(w, q); };
    L.q = qj;
    u128 ratio = (~(u128)0) / qj;  // floor(2^128/q): q odd, never divides 2^128
    L.ratio0 = (u64)ratio; L.ratio1 = (u64)(ratio >> 64);
    L.psi = h_min_primitive_root(qj, n);
    if (!L.psi) { err = "no primitive 2n-th root found"; return false; }
    std::vector<u64> pw(n);
    pw[0] = 1;
    for (uint32_t i = 1; i < n; ++i) pw[i] = h_mulmod(pw[i - 1], L.psi, qj);
    std::vector<W2> seal_tw(n), seal_itw(n);
    for (uint32_t idx = 0; idx < n; ++idx) {
        u64 w = pw[h_bitrev(idx, logn)];
        seal_tw[idx] = shoup(w, qj);
        seal_itw[idx] = shoup(h_invmod(w, qj), qj);
    }
    L.tw.assign(n, W2{0, 0}); L.itw.assign(n, W2{0, 0});
    switch (logn) {
    case 10: ntt_permute_table<10, 4>(seal_tw.data(), L.tw.data()); ntt_permute_table<10, 4>(seal_itw.data(), L.itw.data()); break;
    case 11: ntt_permute_table<11, 4>(seal_tw.data(), L.tw.data()); ntt_permute_table<11, 4>(seal_itw.data(), L.itw.data()); break;
    case 12: ntt_permute_table<12, 4>(seal_tw.data(), L.tw.data()); ntt_permute_table<12, 4>(seal_itw.data(), L.itw.data()); break;
    case 13: ntt_permute_table<13, FHE_LOGE13>(seal_tw.data(), L.tw.data()); ntt_permute_table<13, FHE_LOGE13>(seal_itw.data(), L.itw.data()); break;
    case 14: ntt_permute_table<14, FHE_LOGE14>(seal_tw.data(), L.tw.data()); ntt_permute_table<14, FHE_LOGE14>(seal_itw.data(), L.itw.data()); break;
    default: err = "unsupported ring degree"; return false;
    }
    u64 ninv = h_invmod(n % qj, qj);
    L.itw1 = seal_itw[1].w;
    L.dev.a = h_limb_arith(qj);
    L.dev.ninv = shoup(ninv, qj);
    L.dev.ninv_w1 = shoup(h_mulmod(ninv, L.itw1, qj), qj);
    L.dev.iw1 = seal_itw[1];
    return true;
}

inline bool build_ring(HostRing &R, uint32_t n, uint32_t k, const uint64_t *q, u64 t, std::string &err)
{
    if (n < 1024 || n > 16384 || (n & (n - 1))) { err = "poly_modulus degree must be a power of two in [1024, 16384]"; return false; }
    if (k == 0 || k > 16) { err = "coeff_modulus must have 1..16 primes"; return false; }
    if (t < 2) { err = "plain_modulus must be >= 2"; return false; }
    R = HostRing{};
    R.n = n; R.k = k; R.t = t; R.plain_thr = (t + 1) >> 1;
    while ((1u << R.logn) < n) ++R.logn;
    R.limbs.resize(k);
    for (uint32_t j = 0; j < k; ++j) {
        HostLimb &L = R.limbs[j];
        u64 qj = q[j];
        if (qj >> 60) { err = "coeff_modulus primes must have at most 60 bits"; return false; }  // SEAL 2.3's SmallModulus limit
        if ((qj >> 55) || (qj >> 40) == 0) R.generic = true;  // outside the window of the lazy arithmetic
        if (qj <= t) { err = "coeff_modulus primes must exceed plain_modulus"; return false; }
        if ((qj - 1) % (2ull * n)) { err = "coeff_modulus prime is not 1 mod 2n"; return false; }
        if (!h_is_prime(qj)) { err = "coeff_modulus element is not prime"; return false; }
        for (uint32_t i = 0; i < j; ++i)
            if (q[i] == qj) { err = "coeff_modulus primes must be distinct"; return false; }
        int bits = 64 - __builtin_clzll(qj);
        if (bits > R.max_qbits) R.max_qbits = bits;
    }
    for (uint32_t j = 0; j < k; ++j)
        if (!h_build_limb(R.limbs[j], q[j], n, R.logn, err, R.generic)) return false;
    // q = prod q_j (multi-word), Delta = floor(q/t), q mod t
    std::vector<u64> big(1, 1);
    for (uint32_t j = 0; j < k; ++j) {
        u64 carry = 0;
        for (auto &w : big) {
            u128 p = (u128)w * R.limbs[j].q + carry;
            w = (u64)p; carry = (u64)(p >> 64);
        }
        if (carry) big.push_back(carry);
    }
    std::vector<u64> quo(big.size());
    u64 rem = 0;
    for (int w = (int)big.size() - 1; w >= 0; --w) {
        u128 cur = ((u128)rem << 64) | big[w];
        quo[w] = (u64)(cur / t); rem = (u64)(cur % t);
    }
    R.q_mod_t = rem;
    for (uint32_t j = 0; j < k; ++j) {
        u64 acc = 0;
        for (int w = (int)quo.size() - 1; w >= 0; --w) acc = (u64)((((u128)acc << 64) | quo[w]) % R.limbs[j].q);
        R.limbs[j].delta = acc;
    }
    return true;
}

}  // namespace fhe
