/*
 * fhe_oracle_behz.c -- CPU oracle, part 2 (test infrastructure; see fhe_oracle.h header).
 *
 * PARITY UNPINNED: restates from the published algorithm (Bajard-Eynard-Hasan-Zucca 2016,
 * "A full RNS variant of FV like somewhat homomorphic encryption schemes", as SEAL 2.3 uses it:
 * SURVEY.md Appendix A.6-A.9, [SEAL-recall]) the ciphertext x ciphertext part of the hot path:
 *   Evaluator::multiply / square   call sites homo/fhe_resize.h:174-179,197-198,
 *                                             homo/fhe_decode.h:67-97,147-177,235,239
 *   Evaluator::relinearize         no call site in homo/ (north_star names it; keys only at
 *                                             tests/parameters.cpp:89-90)
 * and the reference op graphs built on them, op by op:
 *   Cubic  fhe_resize.h:143-189    Linear  fhe_resize.h:191-204
 *   homomorphic_sin / cos  fhe_decode.h:48-120, :128-200   approximated_step  :202-242
 * plus the harness-side scheme pieces the reference keeps on the client (keygen, encrypt,
 * the decryption phase c0 + c1 s + c2 s^2 ..., evaluation keys; client_jpeg.cpp:96-116).
 *
 * Every step below is exact modular arithmetic with canonical residues, so the result is a
 * function of the inputs and of the constants only -- any implementation of the same steps
 * (the CUDA kernels use lazy ranges and a different NTT order) must agree bit for bit.
 */
#include "fhe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

#define MAXB (ORC_MAX_LIMBS + 2)

/* SEAL 2.3.1 internal_mods ([SEAL-recall]; recomputed there and in
 * tests/test_oracle_behz.py: the largest primes below 2^61 that are 1 mod 2^18, descending). */
static const uint64_t M_SK = 0x1fffffffffe00001ull;
static const uint64_t AUX_SMALL_MODS[10] = {
    0x1fffffffffb40001ull, 0x1fffffffff500001ull, 0x1fffffffff380001ull, 0x1fffffffff000001ull,
    0x1ffffffffef00001ull, 0x1ffffffffee80001ull, 0x1ffffffffeb40001ull, 0x1ffffffffe780001ull,
    0x1ffffffffe600001ull, 0x1ffffffffe4c0001ull};
#define M_TILDE_BITS 32

/* ------------------------------------------------------------------ one modulus with NTT tables */
typedef struct {
    uint64_t p, r0, r1;
    uint64_t *tw, *twp, *itw, *itwp;
} modtab;

static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) >> 64); }
static inline uint64_t red128(const modtab *m, u128 z)
{
    uint64_t z0 = (uint64_t)z, z1 = (uint64_t)(z >> 64);
    u128 t0 = (u128)z0 * m->r0;
    u128 mid = (u128)z0 * m->r1 + (uint64_t)(t0 >> 64);
    u128 mid2 = (u128)z1 * m->r0 + (uint64_t)mid;
    uint64_t qhat = z1 * m->r1 + (uint64_t)(mid >> 64) + (uint64_t)(mid2 >> 64);
    uint64_t res = z0 - qhat * m->p;
    return res >= m->p ? res - m->p : res;
}
static inline uint64_t mm(const modtab *m, uint64_t a, uint64_t b) { return red128(m, (u128)a * b); }
static inline uint64_t am(uint64_t a, uint64_t b, uint64_t p) { uint64_t s = a + b; return s >= p ? s - p : s; }
static inline uint64_t sm(uint64_t a, uint64_t b, uint64_t p) { return a >= b ? a - b : a + p - b; }
static uint64_t pw(uint64_t b, uint64_t e, uint64_t p)
{
    uint64_t r = 1;
    b %= p;
    while (e) {
        if (e & 1) r = (uint64_t)(((u128)r * b) % p);
        b = (uint64_t)(((u128)b * b) % p);
        e >>= 1;
    }
    return r;
}
static uint64_t inv(uint64_t a, uint64_t p) { return pw(a % p, p - 2, p); } /* p prime */
static uint32_t brev(uint32_t x, uint32_t bits)
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}
static uint64_t min_root(uint64_t q, uint32_t n)
{
    uint64_t root = 0;
    for (uint64_t g = 2; g < 1000; g++) {
        uint64_t r = pw(g, (q - 1) / (2ull * n), q);
        if (pw(r, n, q) == q - 1) {
            root = r;
            break;
        }
    }
    if (!root) return 0;
    uint64_t sq = (uint64_t)(((u128)root * root) % q), cur = root, best = root;
    for (uint32_t i = 0; i < n; i++) {
        if (cur < best) best = cur;
        cur = (uint64_t)(((u128)cur * sq) % q);
    }
    return best;
}
static int modtab_init(modtab *m, uint64_t p, uint32_t n, uint32_t logn, int with_ntt)
{
    memset(m, 0, sizeof(*m));
    m->p = p;
    u128 r = (~(u128)0) / p;
    m->r0 = (uint64_t)r;
    m->r1 = (uint64_t)(r >> 64);
    if (!with_ntt) return 0;
    if ((p - 1) % (2ull * n)) return -1;
    uint64_t psi = min_root(p, n);
    if (!psi) return -1;
    m->tw = (uint64_t *)malloc(sizeof(uint64_t) * n * 4);
    m->twp = m->tw + n;
    m->itw = m->tw + 2 * n;
    m->itwp = m->tw + 3 * n;
    uint64_t *pws = (uint64_t *)malloc(sizeof(uint64_t) * n);
    pws[0] = 1;
    for (uint32_t i = 1; i < n; i++) pws[i] = (uint64_t)(((u128)pws[i - 1] * psi) % p);
    uint64_t inv2 = inv(2, p);
    for (uint32_t idx = 0; idx < n; idx++) {
        uint64_t w = pws[brev(idx, logn)];
        uint64_t wi = (uint64_t)(((u128)inv(w, p) * inv2) % p);
        m->tw[idx] = w;
        m->twp[idx] = (uint64_t)((((u128)w) << 64) / p);
        m->itw[idx] = wi;
        m->itwp[idx] = (uint64_t)((((u128)wi) << 64) / p);
    }
    free(pws);
    return 0;
}
static void modtab_free(modtab *m) { free(m->tw); }

/* Harvey lazy butterflies (needs 4p < 2^64: all moduli here are below 2^62) */
static void ntt_fwd(const modtab *m, uint32_t n, uint64_t *a)
{
    const uint64_t q = m->p, two_q = 2 * q;
    uint32_t t = n >> 1;
    for (uint32_t mm_ = 1; mm_ < n; mm_ <<= 1) {
        for (uint32_t i = 0; i < mm_; i++) {
            uint64_t W = m->tw[mm_ + i], Wp = m->twp[mm_ + i];
            uint64_t *X = a + 2 * i * t, *Y = X + t;
            for (uint32_t s = 0; s < t; s++) {
                uint64_t x = X[s] - (two_q & (uint64_t)(-(int64_t)(X[s] >= two_q)));
                uint64_t Q = Y[s] * W - mulhi64(Wp, Y[s]) * q;
                X[s] = x + Q;
                Y[s] = x + (two_q - Q);
            }
        }
        t >>= 1;
    }
    for (uint32_t i = 0; i < n; i++) {
        uint64_t v = a[i];
        v -= two_q & (uint64_t)(-(int64_t)(v >= two_q));
        v -= q & (uint64_t)(-(int64_t)(v >= q));
        a[i] = v;
    }
}
static void ntt_inv(const modtab *m, uint32_t n, uint64_t *a)
{
    const uint64_t q = m->p, two_q = 2 * q;
    uint32_t t = 1;
    for (uint32_t mm_ = n; mm_ > 1; mm_ >>= 1) {
        uint32_t h = mm_ >> 1, j1 = 0;
        for (uint32_t i = 0; i < h; i++) {
            uint64_t W = m->itw[h + i], Wp = m->itwp[h + i];
            uint64_t *U = a + j1, *V = U + t;
            for (uint32_t s = 0; s < t; s++) {
                uint64_t T = two_q - V[s] + U[s];
                uint64_t cu = U[s] + V[s] - (two_q & (uint64_t)(-(int64_t)((U[s] << 1) >= T)));
                U[s] = (cu + (q & (uint64_t)(-(int64_t)(cu & 1)))) >> 1;
                V[s] = T * W - mulhi64(Wp, T) * q;
            }
            j1 += t << 1;
        }
        t <<= 1;
    }
    for (uint32_t i = 0; i < n; i++) {
        uint64_t v = a[i];
        v -= q & (uint64_t)(-(int64_t)(v >= q));
        a[i] = v;
    }
}

/* ------------------------------------------------------------------ BEHZ context */
struct orc_behz {
    const orc_ctx *c;
    uint32_t n, logn, k, kb, nb; /* kb = |B|, nb = |Bsk| = kb + 1 (m_sk last) */
    uint64_t t;
    modtab qm[ORC_MAX_LIMBS];
    modtab bm[MAXB];
    /* fastbconv_mtilde */
    uint64_t mtilde_qhat_inv[ORC_MAX_LIMBS];     /* m~ (q/q_i)^-1 mod q_i */
    uint64_t qhat_mod_bsk[MAXB][ORC_MAX_LIMBS];  /* (q/q_i) mod bsk_j */
    uint64_t qhat_mod_mtilde[ORC_MAX_LIMBS];     /* (q/q_i) mod 2^32 */
    /* mont_rq */
    uint64_t neg_inv_q_mod_mtilde;               /* -q^-1 mod 2^32 */
    uint64_t q_mod_bsk[MAXB];
    uint64_t inv_mtilde_mod_bsk[MAXB];
    /* fast_floor */
    uint64_t qhat_inv[ORC_MAX_LIMBS];            /* (q/q_i)^-1 mod q_i */
    uint64_t inv_q_mod_bsk[MAXB];
    /* fastbconv_sk */
    uint64_t bhat_inv[MAXB];                     /* (B/b_i)^-1 mod b_i */
    uint64_t bhat_mod_q[ORC_MAX_LIMBS][MAXB];    /* (B/b_i) mod q_j */
    uint64_t bhat_mod_msk[MAXB];
    uint64_t inv_B_mod_msk;
    uint64_t B_mod_q[ORC_MAX_LIMBS];
    uint64_t t_mod_q[ORC_MAX_LIMBS], t_mod_bsk[MAXB];
};

static uint32_t bit_count(uint64_t v)
{
    uint32_t b = 0;
    while (v) {
        b++;
        v >>= 1;
    }
    return b;
}
/* product of all list entries except index skip, modulo p */
static uint64_t prod_except(const uint64_t *list, uint32_t cnt, uint32_t skip, uint64_t p)
{
    uint64_t r = 1 % p;
    for (uint32_t i = 0; i < cnt; i++)
        if (i != skip) r = (uint64_t)(((u128)r * (list[i] % p)) % p);
    return r;
}

orc_behz *orc_behz_create(const orc_ctx *c)
{
    orc_behz *b = (orc_behz *)calloc(1, sizeof(*b));
    b->c = c;
    b->n = orc_n(c);
    b->k = orc_k(c);
    b->t = orc_t(c);
    while ((1u << b->logn) < b->n) b->logn++;
    uint64_t q[ORC_MAX_LIMBS], bs[MAXB];
    uint32_t total_bits = 0;
    for (uint32_t i = 0; i < b->k; i++) {
        q[i] = orc_q(c, i);
        total_bits += bit_count(q[i]);
    }
    /* BaseConverter::generate: |B| = k, one more when the auxiliary base would be too small
     * for K n t q^2 < q B m_sk with 32 bits reserved for K n ([SEAL-recall]). */
    b->kb = b->k;
    if (32 + bit_count(b->t) + total_bits >= 61 * b->k + 61) b->kb++;
    if (b->kb > 10) {
        free(b);
        return NULL;
    }
    b->nb = b->kb + 1;
    for (uint32_t i = 0; i < b->kb; i++) bs[i] = AUX_SMALL_MODS[i];
    bs[b->kb] = M_SK;
    for (uint32_t i = 0; i < b->k; i++)
        if (modtab_init(&b->qm[i], q[i], b->n, b->logn, 1)) goto fail;
    for (uint32_t j = 0; j < b->nb; j++)
        if (modtab_init(&b->bm[j], bs[j], b->n, b->logn, 1)) goto fail;
    const uint64_t MT = 1ull << M_TILDE_BITS;
    for (uint32_t i = 0; i < b->k; i++) {
        uint64_t qhat_i = prod_except(q, b->k, i, q[i]);
        b->qhat_inv[i] = inv(qhat_i, q[i]);
        b->mtilde_qhat_inv[i] = (uint64_t)(((u128)b->qhat_inv[i] * (MT % q[i])) % q[i]);
        b->qhat_mod_mtilde[i] = prod_except(q, b->k, i, MT);
        for (uint32_t j = 0; j < b->nb; j++) b->qhat_mod_bsk[j][i] = prod_except(q, b->k, i, bs[j]);
        b->t_mod_q[i] = b->t % q[i];
        b->B_mod_q[i] = prod_except(bs, b->kb, ~0u, q[i]);
        for (uint32_t l = 0; l < b->kb; l++) b->bhat_mod_q[i][l] = prod_except(bs, b->kb, l, q[i]);
    }
    {
        /* -q^-1 mod 2^32 (Newton iteration on the odd residue) */
        uint64_t qm = prod_except(q, b->k, ~0u, MT), x = 1;
        for (int it = 0; it < 6; it++) x = (x * (2 - qm * x)) & (MT - 1);
        b->neg_inv_q_mod_mtilde = (MT - x) & (MT - 1);
    }
    for (uint32_t j = 0; j < b->nb; j++) {
        b->q_mod_bsk[j] = prod_except(q, b->k, ~0u, bs[j]);
        b->inv_q_mod_bsk[j] = inv(b->q_mod_bsk[j], bs[j]);
        b->inv_mtilde_mod_bsk[j] = inv(MT % bs[j], bs[j]);
        b->t_mod_bsk[j] = b->t % bs[j];
    }
    for (uint32_t l = 0; l < b->kb; l++) {
        b->bhat_inv[l] = inv(prod_except(bs, b->kb, l, bs[l]), bs[l]);
        b->bhat_mod_msk[l] = prod_except(bs, b->kb, l, M_SK);
    }
    b->inv_B_mod_msk = inv(prod_except(bs, b->kb, ~0u, M_SK), M_SK);
    return b;
fail:
    orc_behz_destroy(b);
    return NULL;
}

void orc_behz_destroy(orc_behz *b)
{
    if (!b) return;
    for (uint32_t i = 0; i < ORC_MAX_LIMBS; i++) modtab_free(&b->qm[i]);
    for (uint32_t j = 0; j < MAXB; j++) modtab_free(&b->bm[j]);
    free(b);
}
uint32_t orc_behz_aux_count(const orc_behz *b) { return b->kb; }
uint64_t orc_behz_bsk(const orc_behz *b, uint32_t j) { return b->bm[j].p; }

/* fastbconv_mtilde + mont_rq: one polynomial in base q ([k][n]) ->
 * base Bsk ([nb][n]); the value represented is x + (small multiple of q), free of the
 * q-overflow of the fast conversion. */
void orc_behz_extend(const orc_behz *b, const uint64_t *x, uint64_t *out)
{
    const uint32_t n = b->n, k = b->k;
    const uint64_t MT = 1ull << M_TILDE_BITS;
    uint64_t y[ORC_MAX_LIMBS];
    for (uint32_t c = 0; c < n; c++) {
        uint64_t vm = 0;
        for (uint32_t i = 0; i < k; i++) {
            y[i] = mm(&b->qm[i], x[(size_t)i * n + c], b->mtilde_qhat_inv[i]);
            vm += y[i] * b->qhat_mod_mtilde[i]; /* mod 2^64, reduced mod 2^32 below */
        }
        vm &= MT - 1;
        /* r = [-v q^-1] mod m~, centred in [-m~/2, m~/2) */
        uint64_t r = (vm * b->neg_inv_q_mod_mtilde) & (MT - 1);
        const int neg = r >= (MT >> 1);
        const uint64_t rabs = neg ? MT - r : r;
        for (uint32_t j = 0; j < b->nb; j++) {
            const modtab *m = &b->bm[j];
            u128 acc = 0;
            for (uint32_t i = 0; i < k; i++) acc += (u128)y[i] * b->qhat_mod_bsk[j][i];
            uint64_t v = red128(m, acc);
            uint64_t qr = mm(m, b->q_mod_bsk[j], rabs);
            v = neg ? sm(v, qr, m->p) : am(v, qr, m->p);
            out[(size_t)j * n + c] = mm(m, v, b->inv_mtilde_mod_bsk[j]);
        }
    }
}

/* multiply by t, fast_floor, fastbconv_sk: xq [k][n], xb [nb][n]
 * (coefficient form, tensor results) -> out [k][n] in base q. */
void orc_behz_scale_round(const orc_behz *b, const uint64_t *xq, const uint64_t *xb, uint64_t *out)
{
    const uint32_t n = b->n, k = b->k, kb = b->kb, nb = b->nb;
    const modtab *msk = &b->bm[kb];
    uint64_t y[ORC_MAX_LIMBS], f[MAXB], z[MAXB];
    for (uint32_t c = 0; c < n; c++) {
        for (uint32_t i = 0; i < k; i++) {
            uint64_t tx = mm(&b->qm[i], xq[(size_t)i * n + c], b->t_mod_q[i]);
            y[i] = mm(&b->qm[i], tx, b->qhat_inv[i]);
        }
        /* fast_floor: (t x_bsk - fastbconv_q->bsk(t x_q)) q^-1 */
        for (uint32_t j = 0; j < nb; j++) {
            const modtab *m = &b->bm[j];
            u128 acc = 0;
            for (uint32_t i = 0; i < k; i++) acc += (u128)y[i] * b->qhat_mod_bsk[j][i];
            uint64_t conv = red128(m, acc);
            uint64_t tx = mm(m, xb[(size_t)j * n + c], b->t_mod_bsk[j]);
            f[j] = mm(m, sm(tx, conv, m->p), b->inv_q_mod_bsk[j]);
        }
        /* fastbconv_sk (Shenoy-Kumaresan): B -> q with the overflow alpha read off m_sk */
        for (uint32_t l = 0; l < kb; l++) z[l] = mm(&b->bm[l], f[l], b->bhat_inv[l]);
        u128 acc = 0;
        for (uint32_t l = 0; l < kb; l++) acc += (u128)z[l] * b->bhat_mod_msk[l];
        uint64_t conv_sk = red128(msk, acc);
        uint64_t alpha = mm(msk, sm(conv_sk, f[kb], msk->p), b->inv_B_mod_msk);
        const int neg = alpha > (msk->p >> 1);
        const uint64_t aabs = neg ? msk->p - alpha : alpha;
        for (uint32_t i = 0; i < k; i++) {
            const modtab *m = &b->qm[i];
            u128 a2 = 0;
            for (uint32_t l = 0; l < kb; l++) a2 += (u128)z[l] * b->bhat_mod_q[i][l];
            uint64_t v = red128(m, a2);
            uint64_t ab = mm(m, b->B_mod_q[i], aabs % m->p);
            out[(size_t)i * n + c] = neg ? am(v, ab, m->p) : sm(v, ab, m->p);
        }
    }
}

/* Evaluator::multiply(a, b) -> dest, sizes sa, sb -> sa + sb - 1 (fhe_resize.h:175-179 ...). */
void orc_multiply(const orc_behz *b, const uint64_t *a1, uint32_t s1, const uint64_t *a2, uint32_t s2,
                  uint64_t *dest)
{
    const uint32_t n = b->n, k = b->k, nb = b->nb, sd = s1 + s2 - 1;
    const size_t pq = (size_t)k * n, pb = (size_t)nb * n;
    const int same = (a1 == a2 && s1 == s2);
    uint64_t *q1 = (uint64_t *)malloc(sizeof(uint64_t) * (s1 * (pq + pb) + (same ? 0 : s2 * (pq + pb)) + sd * (pq + pb)));
    uint64_t *b1 = q1 + s1 * pq;
    uint64_t *q2 = same ? q1 : b1 + s1 * pb, *b2 = same ? b1 : q2 + s2 * pq;
    uint64_t *dq = (same ? b1 + s1 * pb : b2 + s2 * pb), *db = dq + sd * pq;
    for (int which = 0; which < (same ? 1 : 2); which++) {
        const uint64_t *src = which ? a2 : a1;
        uint64_t *tq = which ? q2 : q1, *tb = which ? b2 : b1;
        const uint32_t s = which ? s2 : s1;
        memcpy(tq, src, sizeof(uint64_t) * s * pq);
        for (uint32_t p = 0; p < s; p++) {
            orc_behz_extend(b, src + p * pq, tb + p * pb);
            for (uint32_t i = 0; i < k; i++) ntt_fwd(&b->qm[i], n, tq + p * pq + (size_t)i * n);
            for (uint32_t j = 0; j < nb; j++) ntt_fwd(&b->bm[j], n, tb + p * pb + (size_t)j * n);
        }
    }
    memset(dq, 0, sizeof(uint64_t) * sd * (pq + pb));
    for (uint32_t x = 0; x < s1; x++)
        for (uint32_t y = 0; y < s2; y++) {
            for (uint32_t i = 0; i < k; i++) {
                const modtab *m = &b->qm[i];
                const uint64_t *u = q1 + x * pq + (size_t)i * n, *v = q2 + y * pq + (size_t)i * n;
                uint64_t *d = dq + (x + y) * pq + (size_t)i * n;
                for (uint32_t c = 0; c < n; c++) d[c] = am(d[c], mm(m, u[c], v[c]), m->p);
            }
            for (uint32_t j = 0; j < nb; j++) {
                const modtab *m = &b->bm[j];
                const uint64_t *u = b1 + x * pb + (size_t)j * n, *v = b2 + y * pb + (size_t)j * n;
                uint64_t *d = db + (x + y) * pb + (size_t)j * n;
                for (uint32_t c = 0; c < n; c++) d[c] = am(d[c], mm(m, u[c], v[c]), m->p);
            }
        }
    for (uint32_t p = 0; p < sd; p++) {
        for (uint32_t i = 0; i < k; i++) ntt_inv(&b->qm[i], n, dq + p * pq + (size_t)i * n);
        for (uint32_t j = 0; j < nb; j++) ntt_inv(&b->bm[j], n, db + p * pb + (size_t)j * n);
        orc_behz_scale_round(b, dq + p * pq, db + p * pb, dest + p * pq);
    }
    free(q1);
}
/* Evaluator::square (fhe_resize.h:174, fhe_decode.h:67...): the same function of the inputs as
 * multiply(a, a) -- SEAL's size-2 special case computes c0^2, 2 c0 c1, c1^2 in each modulus,
 * which equals the tensor sums exactly. */
void orc_square(const orc_behz *b, const uint64_t *a, uint32_t s, uint64_t *dest)
{
    orc_multiply(b, a, s, a, s, dest);
}

/* ------------------------------------------------------------------ harness: keys, encrypt, phase */
static inline uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } prng;
static uint64_t prng_next(prng *g) { return splitmix64(g->s += 0x9E3779B97F4A7C15ull); }
static double prng_unit(prng *g) { return ((double)(prng_next(g) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
/* clipped normal, sigma 3.19, clip 6 sigma (SEAL default noise_standard_deviation; A.9) */
static int64_t prng_noise(prng *g)
{
    for (;;) {
        double u1 = prng_unit(g), u2 = prng_unit(g);
        double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2) * 3.19;
        double r = nearbyint(z);
        if (fabs(r) <= 19.0) return (int64_t)r;
    }
}
static void small_to_rns(const orc_behz *b, const int64_t *v, uint64_t *out)
{
    for (uint32_t i = 0; i < b->k; i++)
        for (uint32_t c = 0; c < b->n; c++)
            out[(size_t)i * b->n + c] = v[c] >= 0 ? (uint64_t)v[c] : b->qm[i].p - (uint64_t)(-v[c]);
}
/* out = x * y per limb of base q (coefficient form in and out) */
static void poly_mul_q(const orc_behz *b, const uint64_t *x, const uint64_t *y, uint64_t *out)
{
    const uint32_t n = b->n;
    uint64_t *tx = (uint64_t *)malloc(sizeof(uint64_t) * 2 * n), *ty = tx + n;
    for (uint32_t i = 0; i < b->k; i++) {
        const modtab *m = &b->qm[i];
        memcpy(tx, x + (size_t)i * n, sizeof(uint64_t) * n);
        memcpy(ty, y + (size_t)i * n, sizeof(uint64_t) * n);
        ntt_fwd(m, n, tx);
        ntt_fwd(m, n, ty);
        for (uint32_t c = 0; c < n; c++) tx[c] = mm(m, tx[c], ty[c]);
        ntt_inv(m, n, tx);
        memcpy(out + (size_t)i * n, tx, sizeof(uint64_t) * n);
    }
    free(tx);
}
static void poly_add_q(const orc_behz *b, uint64_t *x, const uint64_t *y)
{
    for (uint32_t i = 0; i < b->k; i++)
        for (uint32_t c = 0; c < b->n; c++) {
            size_t o = (size_t)i * b->n + c;
            x[o] = am(x[o], y[o], b->qm[i].p);
        }
}
static void poly_uniform_q(const orc_behz *b, prng *g, uint64_t *out)
{
    for (uint32_t i = 0; i < b->k; i++)
        for (uint32_t c = 0; c < b->n; c++) out[(size_t)i * b->n + c] = mulhi64(prng_next(g), b->qm[i].p);
}
static void poly_noise_q(const orc_behz *b, prng *g, uint64_t *out)
{
    int64_t *e = (int64_t *)malloc(sizeof(int64_t) * b->n);
    for (uint32_t c = 0; c < b->n; c++) e[c] = prng_noise(g);
    small_to_rns(b, e, out);
    free(e);
}
static void poly_ternary_q(const orc_behz *b, prng *g, uint64_t *out)
{
    int64_t *e = (int64_t *)malloc(sizeof(int64_t) * b->n);
    for (uint32_t c = 0; c < b->n; c++) e[c] = (int64_t)(prng_next(g) % 3) - 1;
    small_to_rns(b, e, out);
    free(e);
}

/* KeyGenerator (client_jpeg.cpp:96-103): sk ternary [k][n]; pk = (-(a s + e), a) [2][k][n],
 * both in coefficient form. */
void orc_keygen(const orc_behz *b, uint64_t seed, uint64_t *sk, uint64_t *pk)
{
    prng g = {seed};
    const size_t pq = (size_t)b->k * b->n;
    poly_ternary_q(b, &g, sk);
    uint64_t *e = (uint64_t *)malloc(sizeof(uint64_t) * pq);
    poly_uniform_q(b, &g, pk + pq);
    poly_noise_q(b, &g, e);
    poly_mul_q(b, pk + pq, sk, pk);
    poly_add_q(b, pk, e);
    for (uint32_t i = 0; i < b->k; i++)
        for (uint32_t c = 0; c < b->n; c++) {
            size_t o = (size_t)i * b->n + c;
            pk[o] = pk[o] ? b->qm[i].p - pk[o] : 0;
        }
    free(e);
}
/* Encryptor::encrypt (client_jpeg.cpp:147-162; server-side uses at fhe_resize.h:230-266,
 * fhe_decode.h:54,134): ct = (Delta m + pk0 u + e1, pk1 u + e2). */
void orc_encrypt(const orc_behz *b, const uint64_t *pk, const uint64_t *plain, uint32_t plain_len,
                 uint64_t seed, uint64_t *ct)
{
    prng g = {seed ^ 0xE7C3A1D5B2F49687ull};
    const size_t pq = (size_t)b->k * b->n;
    uint64_t *u = (uint64_t *)malloc(sizeof(uint64_t) * 2 * pq), *e = u + pq;
    poly_ternary_q(b, &g, u);
    for (int p = 0; p < 2; p++) {
        poly_mul_q(b, pk + p * pq, u, ct + p * pq);
        poly_noise_q(b, &g, e);
        poly_add_q(b, ct + p * pq, e);
    }
    orc_add_plain(b->c, ct, plain, plain_len);
    free(u);
}
/* Decryptor phase: c0 + c1 s + c2 s^2 + ... per limb ([k][n]); the harness (oracle_py.py)
 * finishes with exact big-integer rounding round(t x / q) mod t. */
void orc_phase(const orc_behz *b, const uint64_t *sk, const uint64_t *ct, uint32_t size, uint64_t *out)
{
    const size_t pq = (size_t)b->k * b->n;
    uint64_t *sp = (uint64_t *)malloc(sizeof(uint64_t) * 2 * pq), *tmp = sp + pq;
    memcpy(out, ct, sizeof(uint64_t) * pq);
    memcpy(sp, sk, sizeof(uint64_t) * pq);
    for (uint32_t p = 1; p < size; p++) {
        poly_mul_q(b, ct + p * pq, sp, tmp);
        poly_add_q(b, out, tmp);
        if (p + 1 < size) poly_mul_q(b, sp, sk, sp);
    }
    free(sp);
}

/* KeyGenerator::generate_evaluation_keys(dbc, count) (tests/parameters.cpp:89-90):
 * for r < count (power s^(2+r)), limb i, digit l < ceil(bits(q_i)/dbc): an RLWE pair
 * (-(a s + e) + s^(2+r) w^l (q/q_i) [(q/q_i)^-1]_{q_i}, a), w = 2^dbc.  The CRT idempotent
 * (q/q_i)[(q/q_i)^-1] is 1 mod q_i and 0 mod the other primes, so only limb i of the message
 * term is non-zero.  Layout [r][i][l][2][k][n], COEFFICIENT form (SEAL keeps them in NTT form;
 * the NTT order is an internal convention of each implementation).  Returns the number of
 * digits per limb through digits_out (max over limbs, missing digits are zero keys). */
uint32_t orc_evk_digits(const orc_behz *b, uint32_t dbc)
{
    uint32_t d = 0;
    for (uint32_t i = 0; i < b->k; i++) {
        uint32_t di = (bit_count(b->qm[i].p) + dbc - 1) / dbc;
        if (di > d) d = di;
    }
    return d;
}
void orc_evk_gen(const orc_behz *b, const uint64_t *sk, uint32_t dbc, uint32_t count, uint64_t seed,
                 uint64_t *evk)
{
    prng g = {seed ^ 0x1F83D9ABFB41BD6Bull};
    const uint32_t n = b->n, k = b->k, nd = orc_evk_digits(b, dbc);
    const size_t pq = (size_t)k * n;
    uint64_t *sp = (uint64_t *)malloc(sizeof(uint64_t) * 2 * pq), *e = sp + pq;
    poly_mul_q(b, sk, sk, sp); /* s^2 */
    for (uint32_t r = 0; r < count; r++) {
        for (uint32_t i = 0; i < k; i++)
            for (uint32_t l = 0; l < nd; l++) {
                uint64_t *k0 = evk + ((((size_t)r * k + i) * nd + l) * 2) * pq, *k1 = k0 + pq;
                poly_uniform_q(b, &g, k1);
                poly_noise_q(b, &g, e);
                poly_mul_q(b, k1, sk, k0);
                poly_add_q(b, k0, e);
                const modtab *m = &b->qm[i];
                const uint64_t wl = pw(2, (uint64_t)dbc * l, m->p);
                for (uint32_t j = 0; j < k; j++)
                    for (uint32_t c = 0; c < n; c++) {
                        size_t o = (size_t)j * n + c;
                        uint64_t v = k0[o] ? b->qm[j].p - k0[o] : 0;
                        if (j == i) v = am(v, mm(m, sp[o], wl), m->p);
                        k0[o] = v;
                    }
            }
        if (r + 1 < count) poly_mul_q(b, sp, sk, sp);
    }
    free(sp);
}
/* Evaluator::relinearize(ct, evk) down to size 2: one step removes the last poly
 * c_last: digit polys d_{i,l} = (c_last mod q_i >> l dbc) & (w - 1), small non-negative
 * integers read in every limb j; c0 += sum d evk0, c1 += sum d evk1.  key r serves size 3 + r. */
void orc_relinearize(const orc_behz *b, const uint64_t *ct, uint32_t size, const uint64_t *evk,
                     uint32_t dbc, uint32_t evk_count, uint64_t *dest)
{
    const uint32_t n = b->n, k = b->k, nd = orc_evk_digits(b, dbc);
    const size_t pq = (size_t)k * n;
    uint64_t *work = (uint64_t *)malloc(sizeof(uint64_t) * (size * pq + 3 * pq));
    uint64_t *d = work + size * pq, *acc = d + pq, *key = acc + pq;
    memcpy(work, ct, sizeof(uint64_t) * size * pq);
    const uint64_t mask = dbc >= 64 ? ~0ull : ((1ull << dbc) - 1);
    while (size > 2) {
        const uint32_t r = size - 3;
        if (r >= evk_count) break;
        const uint64_t *last = work + (size - 1) * pq;
        for (uint32_t i = 0; i < k; i++)
            for (uint32_t l = 0; l < nd; l++) {
                const uint64_t *k0 = evk + ((((size_t)r * k + i) * nd + l) * 2) * pq;
                for (uint32_t j = 0; j < k; j++)
                    for (uint32_t c = 0; c < n; c++)
                        d[(size_t)j * n + c] = (last[(size_t)i * n + c] >> (l * dbc)) & mask;
                for (int p = 0; p < 2; p++) {
                    memcpy(key, k0 + p * pq, sizeof(uint64_t) * pq);
                    poly_mul_q(b, d, key, acc);
                    poly_add_q(b, work + p * pq, acc);
                }
            }
        size--;
    }
    memcpy(dest, work, sizeof(uint64_t) * 2 * pq);
    free(work);
}

/* ------------------------------------------------------------------ variable-size ciphertext helpers */
typedef struct {
    uint32_t size;
    uint64_t *v; /* capacity ORC_MAX_CT_SIZE polys */
} oct;
#define ORC_MAX_CT_SIZE 24
typedef struct {
    const orc_behz *b;
    uint32_t ic, fc;
    size_t pq;
    uint64_t *plain;
} bg;

static oct oct_new(const bg *g)
{
    oct o = {0, (uint64_t *)calloc((size_t)ORC_MAX_CT_SIZE * g->pq, sizeof(uint64_t))};
    return o;
}
static void oct_set(const bg *g, oct *dst, const uint64_t *src, uint32_t size)
{
    dst->size = size;
    memcpy(dst->v, src, sizeof(uint64_t) * size * g->pq);
}
static void oct_copy(const bg *g, oct *dst, const oct *src) { oct_set(g, dst, src->v, src->size); }
/* Evaluator::add / sub with unequal sizes: result size = max, missing polys are 0 */
static void oct_addsub(const bg *g, oct *a, const oct *b, int sub)
{
    const orc_behz *B = g->b;
    uint32_t mx = a->size > b->size ? a->size : b->size;
    if (a->size < mx) memset(a->v + a->size * g->pq, 0, sizeof(uint64_t) * (mx - a->size) * g->pq);
    for (uint32_t p = 0; p < b->size; p++)
        for (uint32_t i = 0; i < B->k; i++) {
            const uint64_t q = B->qm[i].p;
            uint64_t *x = a->v + p * g->pq + (size_t)i * B->n;
            const uint64_t *y = b->v + p * g->pq + (size_t)i * B->n;
            for (uint32_t c = 0; c < B->n; c++) x[c] = sub ? sm(x[c], y[c], q) : am(x[c], y[c], q);
        }
    a->size = mx;
}
static void oct_mulp(const bg *g, oct *a, double v)
{
    orc_frac_encode(g->b->c, v, g->ic, g->fc, g->plain);
    orc_multiply_plain(g->b->c, a->v, a->size, g->plain, g->b->n);
}
static void oct_addp(const bg *g, oct *a, double v)
{
    orc_frac_encode(g->b->c, v, g->ic, g->fc, g->plain);
    orc_add_plain(g->b->c, a->v, g->plain, g->b->n);
}
static void oct_mul(const bg *g, oct *a, const oct *b)
{
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * (a->size + b->size - 1) * g->pq);
    orc_multiply(g->b, a->v, a->size, b->v, b->size, tmp);
    a->size = a->size + b->size - 1;
    memcpy(a->v, tmp, sizeof(uint64_t) * a->size * g->pq);
    free(tmp);
}
static void oct_square(const bg *g, oct *a) { oct_mul(g, a, a); }

/* Cubic, fhe_resize.h:143-189, op by op (including t3 = t*t, Appendix C.1).
 * in[0..3] = A, B, C, D with sizes sz[0..3]; t of size st; result size returned. */
uint32_t orc_cubic(const orc_behz *b, const uint64_t *const in[4], const uint32_t sz[4],
                   const uint64_t *t, uint32_t st, uint32_t ic, uint32_t fc, uint64_t *out)
{
    bg g = {b, ic, fc, (size_t)b->k * b->n, (uint64_t *)malloc(sizeof(uint64_t) * b->n)};
    oct A = oct_new(&g), B = oct_new(&g), C = oct_new(&g), D = oct_new(&g), T = oct_new(&g);
    oct a = oct_new(&g), bb = oct_new(&g), c = oct_new(&g), x = oct_new(&g), t2 = oct_new(&g), t3 = oct_new(&g);
    oct_set(&g, &A, in[0], sz[0]);
    oct_set(&g, &B, in[1], sz[1]);
    oct_set(&g, &C, in[2], sz[2]);
    oct_set(&g, &D, in[3], sz[3]);
    oct_set(&g, &T, t, st);
    /* a = 3B - A - 3C + D  (:149-156) */
    oct_copy(&g, &a, &B);
    oct_mulp(&g, &a, 3);
    oct_addsub(&g, &a, &A, 1);
    oct_copy(&g, &x, &C);
    oct_mulp(&g, &x, 3);
    oct_addsub(&g, &a, &x, 1);
    oct_addsub(&g, &a, &D, 0);
    /* b = 2A - 5B + 4C - D  (:158-165) */
    oct_copy(&g, &bb, &A);
    oct_mulp(&g, &bb, 2);
    oct_copy(&g, &x, &B);
    oct_mulp(&g, &x, 5);
    oct_addsub(&g, &bb, &x, 1);
    oct_copy(&g, &x, &C);
    oct_mulp(&g, &x, 4);
    oct_addsub(&g, &bb, &x, 0);
    oct_addsub(&g, &bb, &D, 1);
    /* c = C - A  (:167-170) */
    oct_copy(&g, &c, &C);
    oct_addsub(&g, &c, &A, 1);
    /* t2 = square(t); t3 = t * t  (:174-175) */
    oct_copy(&g, &t2, &T);
    oct_square(&g, &t2);
    oct_copy(&g, &t3, &T);
    oct_mul(&g, &t3, &T);
    oct_mul(&g, &a, &t3);
    oct_mul(&g, &bb, &t2);
    oct_mul(&g, &c, &T);
    oct_addsub(&g, &a, &bb, 0);
    oct_addsub(&g, &a, &c, 0);
    oct_mulp(&g, &a, 0.5);
    oct_addsub(&g, &a, &B, 0); /* d = B */
    memcpy(out, a.v, sizeof(uint64_t) * a.size * g.pq);
    uint32_t rs = a.size;
    free(A.v); free(B.v); free(C.v); free(D.v); free(T.v);
    free(a.v); free(bb.v); free(c.v); free(x.v); free(t2.v); free(t3.v);
    free(g.plain);
    return rs;
}

/* Linear, fhe_resize.h:191-204: (1 - t) A + t B. */
uint32_t orc_linear(const orc_behz *b, const uint64_t *A, uint32_t sA, const uint64_t *Bc, uint32_t sB,
                    const uint64_t *t, uint32_t st, uint32_t ic, uint32_t fc, uint64_t *out)
{
    bg g = {b, ic, fc, (size_t)b->k * b->n, (uint64_t *)malloc(sizeof(uint64_t) * b->n)};
    oct x = oct_new(&g), y = oct_new(&g), a = oct_new(&g), bb = oct_new(&g), T = oct_new(&g);
    oct_set(&g, &a, A, sA);
    oct_set(&g, &bb, Bc, sB);
    oct_set(&g, &T, t, st);
    oct_copy(&g, &x, &T);
    orc_negate(b->c, x.v, x.size);
    oct_addp(&g, &x, 1.0);
    oct_mul(&g, &x, &a);
    oct_copy(&g, &y, &bb);
    oct_mul(&g, &y, &T);
    oct_addsub(&g, &x, &y, 0);
    memcpy(out, x.v, sizeof(uint64_t) * x.size * g.pq);
    uint32_t rs = x.size;
    free(x.v); free(y.v); free(a.v); free(bb.v); free(T.v);
    free(g.plain);
    return rs;
}

/* homomorphic_sin / homomorphic_cos, fhe_decode.h:48-120 / :128-200.  Both expand around
 * 3 pi / 2 (Appendix C.3); they differ only in the signs of the Taylor coefficients.
 * zero = the fresh server-side encryption of encode(0.0) (:54, :134; an INPUT here, C.4). */
static void taylor_sincos(const bg *g, const oct *x, const oct *zero, int is_cos, oct *res)
{
    const double sgn = is_cos ? -1.0 : 1.0;
    oct sh = oct_new(g), p = oct_new(g);
    oct_copy(g, res, zero);
    oct_copy(g, &sh, x);
    oct_addp(g, &sh, -3 * M_PI / 2.0);
    oct_addp(g, res, sgn * -1.0); /* :113 / :193; order of the exact additions does not matter */
    /* 2nd order */
    oct_copy(g, &p, &sh);
    oct_square(g, &p);
    oct_mulp(g, &p, sgn * 0.5);
    oct_addsub(g, res, &p, 0);
    /* 4th */
    oct_copy(g, &p, &sh);
    oct_square(g, &p);
    oct_square(g, &p);
    oct_mulp(g, &p, sgn * (-1.0 / 24.0));
    oct_addsub(g, res, &p, 0);
    /* 6th */
    oct_copy(g, &p, &sh);
    oct_square(g, &p);
    oct_square(g, &p);
    oct_mul(g, &p, &sh);
    oct_mul(g, &p, &sh);
    oct_mulp(g, &p, sgn * (1.0 / 720.0));
    oct_addsub(g, res, &p, 0);
    /* 8th */
    oct_copy(g, &p, &sh);
    oct_square(g, &p);
    oct_square(g, &p);
    oct_square(g, &p);
    oct_mulp(g, &p, sgn * (-1.0 / 40320.0));
    oct_addsub(g, res, &p, 0);
    /* 10th */
    oct_copy(g, &p, &sh);
    oct_square(g, &p);
    oct_square(g, &p);
    oct_square(g, &p);
    oct_mul(g, &p, &sh);
    oct_mul(g, &p, &sh);
    oct_mulp(g, &p, sgn * (1.0 / 3628800.0));
    oct_addsub(g, res, &p, 0);
    free(sh.v);
    free(p.v);
}
uint32_t orc_homomorphic_sincos(const orc_behz *b, const uint64_t *x, uint32_t sx, const uint64_t *zero,
                                int is_cos, uint32_t ic, uint32_t fc, uint64_t *out)
{
    bg g = {b, ic, fc, (size_t)b->k * b->n, (uint64_t *)malloc(sizeof(uint64_t) * b->n)};
    oct X = oct_new(&g), Z = oct_new(&g), R = oct_new(&g);
    oct_set(&g, &X, x, sx);
    oct_set(&g, &Z, zero, 2);
    taylor_sincos(&g, &X, &Z, is_cos, &R);
    memcpy(out, R.v, sizeof(uint64_t) * R.size * g.pq);
    uint32_t rs = R.size;
    free(X.v); free(Z.v); free(R.v);
    free(g.plain);
    return rs;
}

/* approximated_step (homomorphic overload), fhe_decode.h:202-242, including the add_plain of
 * encode(i) to `offset` INSIDE the j loop (Appendix C.2).  amplitude, index, count: size 2.
 * zeros: 2 * degree * npos fresh encryptions of 0, consumed in call order (sin then cos).
 * run_out: npos ciphertexts of the returned size each. */
uint32_t orc_approximated_step(const orc_behz *b, const uint64_t *amplitude, const uint64_t *index,
                               const uint64_t *count, int order, int degree, double delta, int npos,
                               const uint64_t *zeros, uint32_t ic, uint32_t fc, uint64_t *run_out)
{
    bg g = {b, ic, fc, (size_t)b->k * b->n, (uint64_t *)malloc(sizeof(uint64_t) * b->n)};
    oct bb = oct_new(&g), off = oct_new(&g), c = oct_new(&g), sa = oct_new(&g), ca = oct_new(&g);
    oct sf = oct_new(&g), cf = oct_new(&g), amp = oct_new(&g), z = oct_new(&g);
    oct_set(&g, &amp, amplitude, 2);
    oct_set(&g, &bb, count, 2);
    oct_mulp(&g, &bb, 0.5);
    oct_set(&g, &off, index, 2);
    oct_addsub(&g, &off, &bb, 0);
    oct_addp(&g, &off, -0.5);
    orc_negate(b->c, off.v, off.size);
    oct_addp(&g, &bb, delta - 0.5);
    uint32_t rs = 0;
    size_t zi = 0;
    for (int i = 0; i < npos; i++) {
        oct_copy(&g, &c, &bb);
        oct_mulp(&g, &c, 1.0 / ((double)order));
        for (int j = 1; j <= degree; j++) {
            double arg_factor = ((float)j) * M_PI / ((double)order);
            oct_copy(&g, &sa, &bb);
            oct_mulp(&g, &sa, arg_factor);
            oct_copy(&g, &ca, &off);
            if (i != 0) oct_addp(&g, &off, (double)i); /* encode(0) is the zero plaintext: add_plain of it is the identity */
            oct_mulp(&g, &ca, arg_factor);
            oct_set(&g, &z, zeros + (zi++) * 2 * g.pq, 2);
            taylor_sincos(&g, &sa, &z, 0, &sf);
            oct_set(&g, &z, zeros + (zi++) * 2 * g.pq, 2);
            taylor_sincos(&g, &ca, &z, 1, &cf);
            oct_mul(&g, &sf, &cf);
            oct_mulp(&g, &sf, 2.0 / (M_PI * ((float)j)));
            oct_addsub(&g, &c, &sf, 0);
        }
        oct_mul(&g, &c, &amp);
        rs = c.size;
        memcpy(run_out + (size_t)i * rs * g.pq, c.v, sizeof(uint64_t) * rs * g.pq);
    }
    free(bb.v); free(off.v); free(c.v); free(sa.v); free(ca.v); free(sf.v); free(cf.v); free(amp.v); free(z.v);
    free(g.plain);
    return rs;
}
