/*
 * fhe_oracle.c -- CPU oracle (test infrastructure; see fhe_oracle.h header).
 * PARITY UNPINNED: SEAL 2.3 sources are absent from /root/reference; this is a
 * restatement of the published algorithm, anchored on
 * the reference call sites cited at each function.
 */
#include "fhe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

typedef unsigned __int128 u128;

struct orc_ctx {
    uint32_t n, logn, k;
    uint64_t t;
    uint64_t q[ORC_MAX_LIMBS];
    uint64_t ratio0[ORC_MAX_LIMBS], ratio1[ORC_MAX_LIMBS]; /* floor(2^128/q) */
    uint64_t psi[ORC_MAX_LIMBS];
    uint64_t delta[ORC_MAX_LIMBS]; /* floor(q/t) mod q_j */
    uint64_t q_mod_t;
    uint64_t plain_upper_half_threshold; /* (t+1)/2 */
    /* tw[j][idx] = psi^bitrev(idx) (SEAL "root powers", scrambled order), Shoup-scaled copy,
     * itw = inverse of tw divided by two. */
    uint64_t *tw[ORC_MAX_LIMBS], *twp[ORC_MAX_LIMBS];
    uint64_t *itw[ORC_MAX_LIMBS], *itwp[ORC_MAX_LIMBS];
};

/* ------------------------------------------------------------------ arithmetic */

static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) >> 64); }

/* Barrett 128 -> 64 (SEAL util/uintarithsmallmod barrett_reduce_128, restated). */
static inline uint64_t barrett128(uint64_t z0, uint64_t z1, uint64_t q, uint64_t r0, uint64_t r1)
{
    u128 t0 = (u128)z0 * r0;
    u128 mid = (u128)z0 * r1 + (uint64_t)(t0 >> 64);
    u128 mid2 = (u128)z1 * r0 + (uint64_t)mid;
    uint64_t qhat = z1 * r1 + (uint64_t)(mid >> 64) + (uint64_t)(mid2 >> 64);
    uint64_t res = z0 - qhat * q;
    return res >= q ? res - q : res;
}

static inline uint64_t mulmod(const orc_ctx *c, uint32_t j, uint64_t a, uint64_t b)
{
    u128 z = (u128)a * b;
    return barrett128((uint64_t)z, (uint64_t)(z >> 64), c->q[j], c->ratio0[j], c->ratio1[j]);
}
static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t q)
{
    uint64_t s = a + b;
    return s >= q ? s - q : s;
}
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t q) { return a >= b ? a - b : a + q - b; }

static uint64_t powmod_slow(uint64_t b, uint64_t e, uint64_t q)
{
    uint64_t r = 1;
    b %= q;
    while (e) {
        if (e & 1) r = (uint64_t)(((u128)r * b) % q);
        b = (uint64_t)(((u128)b * b) % q);
        e >>= 1;
    }
    return r;
}
static uint64_t invmod_slow(uint64_t a, uint64_t q) { return powmod_slow(a, q - 2, q); }
static inline uint64_t shoup(uint64_t w, uint64_t q) { return (uint64_t)((((u128)w) << 64) / q); }

static uint32_t bitrev(uint32_t x, uint32_t bits)
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

/* SEAL util try_minimal_primitive_root, restated: a primitive 2n-th root of unity, then
 * the smallest among its odd powers (all primitive 2n-th roots). */
static uint64_t minimal_primitive_root(uint64_t q, uint32_t n)
{
    uint64_t two_n = 2ull * n, root = 0;
    for (uint64_t g = 2; g < 1000; g++) {
        uint64_t r = powmod_slow(g, (q - 1) / two_n, q);
        if (powmod_slow(r, n, q) == q - 1) {
            root = r;
            break;
        }
    }
    if (!root) return 0;
    uint64_t gen_sq = (uint64_t)(((u128)root * root) % q), cur = root, best = root;
    for (uint32_t i = 0; i < n; i++) {
        if (cur < best) best = cur;
        cur = (uint64_t)(((u128)cur * gen_sq) % q);
    }
    return best;
}

/* ------------------------------------------------------------------ context */

uint32_t orc_default_coeff_modulus(uint32_t n, uint64_t *q_out)
{
    /* SEAL 2.3.1 small_mods_55bit prefix; SURVEY.md section 0.1 (checked there: the largest
     * primes below 2^55 that are 1 mod 2^18, descending). */
    static const uint64_t P[8] = {0x7fffffff380001ull, 0x7ffffffef00001ull, 0x7ffffffeac0001ull,
                                  0x7ffffffe700001ull, 0x7ffffffe600001ull, 0x7ffffffe4c0001ull,
                                  0x7ffffffdec0001ull, 0x7ffffffdd00001ull};
    uint32_t k;
    switch (n) {
    case 2048: q_out[0] = 0x3fffffff000001ull; return 1;
    case 4096: k = 2; break;
    case 8192: k = 4; break;
    case 16384: k = 8; break;
    default: return 0;
    }
    for (uint32_t i = 0; i < k; i++) q_out[i] = P[i];
    return k;
}

orc_ctx *orc_ctx_create(uint32_t n, uint32_t k, const uint64_t *q, uint64_t t)
{
    if (k == 0 || k > ORC_MAX_LIMBS || n < 2 || (n & (n - 1)) || t < 2) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(*c));
    c->n = n;
    c->k = k;
    c->t = t;
    while ((1u << c->logn) < n) c->logn++;
    c->plain_upper_half_threshold = (t + 1) >> 1;
    for (uint32_t j = 0; j < k; j++) {
        uint64_t qj = q[j];
        if (qj >> 62 || (qj - 1) % (2ull * n) != 0 || qj <= t) {
            orc_ctx_destroy(c);
            return NULL;
        }
        c->q[j] = qj;
        /* floor(2^128 / q): 2^128 = (2^128-1) + 1 and q odd>1 never divides 2^128 */
        u128 all = ~(u128)0;
        u128 r = all / qj;
        c->ratio0[j] = (uint64_t)r;
        c->ratio1[j] = (uint64_t)(r >> 64);
        c->psi[j] = minimal_primitive_root(qj, n);
        if (!c->psi[j]) {
            orc_ctx_destroy(c);
            return NULL;
        }
        c->tw[j] = (uint64_t *)malloc(sizeof(uint64_t) * n);
        c->twp[j] = (uint64_t *)malloc(sizeof(uint64_t) * n);
        c->itw[j] = (uint64_t *)malloc(sizeof(uint64_t) * n);
        c->itwp[j] = (uint64_t *)malloc(sizeof(uint64_t) * n);
        uint64_t inv2 = invmod_slow(2, qj);
        uint64_t *pw = (uint64_t *)malloc(sizeof(uint64_t) * n);
        pw[0] = 1;
        for (uint32_t i = 1; i < n; i++) pw[i] = (uint64_t)(((u128)pw[i - 1] * c->psi[j]) % qj);
        for (uint32_t idx = 0; idx < n; idx++) {
            uint64_t w = pw[bitrev(idx, c->logn)];
            uint64_t wi = (uint64_t)(((u128)invmod_slow(w, qj) * inv2) % qj);
            c->tw[j][idx] = w;
            c->twp[j][idx] = shoup(w, qj);
            c->itw[j][idx] = wi;
            c->itwp[j][idx] = shoup(wi, qj);
        }
        free(pw);
    }
    /* Delta = floor(q/t) and q mod t with q = prod q_j: schoolbook multi-word product, then
     * single-word long division (SEAL context "coeff_div_plain_modulus", "upper_half_increment"). */
    uint64_t big[ORC_MAX_LIMBS + 1];
    memset(big, 0, sizeof(big));
    big[0] = 1;
    uint32_t words = 1;
    for (uint32_t j = 0; j < k; j++) {
        uint64_t carry = 0;
        for (uint32_t w = 0; w < words; w++) {
            u128 p = (u128)big[w] * c->q[j] + carry;
            big[w] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        if (carry) big[words++] = carry;
    }
    uint64_t quo[ORC_MAX_LIMBS + 1];
    uint64_t rem = 0;
    for (int w = (int)words - 1; w >= 0; w--) {
        u128 cur = ((u128)rem << 64) | big[w];
        quo[w] = (uint64_t)(cur / t);
        rem = (uint64_t)(cur % t);
    }
    c->q_mod_t = rem;
    for (uint32_t j = 0; j < k; j++) {
        uint64_t acc = 0;
        for (int w = (int)words - 1; w >= 0; w--) {
            u128 cur = ((u128)acc << 64) | quo[w];
            acc = (uint64_t)(cur % c->q[j]);
        }
        c->delta[j] = acc;
    }
    return c;
}

void orc_ctx_destroy(orc_ctx *c)
{
    if (!c) return;
    for (uint32_t j = 0; j < ORC_MAX_LIMBS; j++) {
        free(c->tw[j]);
        free(c->twp[j]);
        free(c->itw[j]);
        free(c->itwp[j]);
    }
    free(c);
}

uint32_t orc_n(const orc_ctx *c) { return c->n; }
uint32_t orc_k(const orc_ctx *c) { return c->k; }
uint64_t orc_q(const orc_ctx *c, uint32_t j) { return c->q[j]; }
uint64_t orc_t(const orc_ctx *c) { return c->t; }
uint64_t orc_psi(const orc_ctx *c, uint32_t j) { return c->psi[j]; }
uint64_t orc_delta_mod(const orc_ctx *c, uint32_t j) { return c->delta[j]; }
uint64_t orc_q_mod_t(const orc_ctx *c) { return c->q_mod_t; }

/* ------------------------------------------------------------------ NTT
 * Harvey lazy butterflies (SEAL util/smallntt ntt_negacyclic_harvey_lazy and
 * inverse_ntt_negacyclic_harvey, restated). */

void orc_ntt_forward(const orc_ctx *c, uint32_t j, uint64_t *a)
{
    const uint64_t q = c->q[j], two_q = 2 * q;
    const uint64_t *tw = c->tw[j], *twp = c->twp[j];
    uint32_t n = c->n, t = n >> 1;
    for (uint32_t m = 1; m < n; m <<= 1) {
        for (uint32_t i = 0; i < m; i++) {
            uint64_t W = tw[m + i], Wp = twp[m + i];
            uint64_t *X = a + 2 * i * t, *Y = X + t;
            for (uint32_t s = 0; s < t; s++) {
                uint64_t x = X[s] - (two_q & (uint64_t)(-(int64_t)(X[s] >= two_q)));
                uint64_t Q = mulhi64(Wp, Y[s]);
                Q = Y[s] * W - Q * q;
                X[s] = x + Q;
                Y[s] = x + (two_q - Q);
            }
        }
        t >>= 1;
    }
    for (uint32_t i = 0; i < n; i++) { /* [0,4q) -> [0,q) */
        uint64_t v = a[i];
        v -= two_q & (uint64_t)(-(int64_t)(v >= two_q));
        v -= q & (uint64_t)(-(int64_t)(v >= q));
        a[i] = v;
    }
}

void orc_ntt_inverse(const orc_ctx *c, uint32_t j, uint64_t *a)
{
    const uint64_t q = c->q[j], two_q = 2 * q;
    const uint64_t *tw = c->itw[j], *twp = c->itwp[j];
    uint32_t n = c->n, t = 1;
    for (uint32_t m = n; m > 1; m >>= 1) {
        uint32_t h = m >> 1, j1 = 0;
        for (uint32_t i = 0; i < h; i++) {
            uint64_t W = tw[h + i], Wp = twp[h + i];
            uint64_t *U = a + j1, *V = U + t;
            for (uint32_t s = 0; s < t; s++) {
                uint64_t T = two_q - V[s] + U[s];
                uint64_t cu = U[s] + V[s] - (two_q & (uint64_t)(-(int64_t)((U[s] << 1) >= T)));
                U[s] = (cu + (q & (uint64_t)(-(int64_t)(cu & 1)))) >> 1;
                uint64_t Q = mulhi64(Wp, T);
                V[s] = T * W - Q * q;
            }
            j1 += t << 1;
        }
        t <<= 1;
    }
    for (uint32_t i = 0; i < n; i++) { /* [0,2q) -> [0,q) */
        uint64_t v = a[i];
        v -= q & (uint64_t)(-(int64_t)(v >= q));
        a[i] = v;
    }
}

void orc_negacyclic_mul(const orc_ctx *c, uint32_t j, const uint64_t *a, const uint64_t *b,
                        uint64_t *out)
{
    uint32_t n = c->n;
    uint64_t *ta = (uint64_t *)malloc(sizeof(uint64_t) * n * 2), *tb = ta + n;
    memcpy(ta, a, sizeof(uint64_t) * n);
    memcpy(tb, b, sizeof(uint64_t) * n);
    orc_ntt_forward(c, j, ta);
    orc_ntt_forward(c, j, tb);
    for (uint32_t i = 0; i < n; i++) ta[i] = mulmod(c, j, ta[i], tb[i]);
    orc_ntt_inverse(c, j, ta);
    memcpy(out, ta, sizeof(uint64_t) * n);
    free(ta);
}

/* ------------------------------------------------------------------ Evaluator subset */

static inline size_t ct_words(const orc_ctx *c, uint32_t size) { return (size_t)size * c->k * c->n; }

/* Evaluator::add; call sites e.g. fhe_image.h:207. */
void orc_add(const orc_ctx *c, uint64_t *a, const uint64_t *b, uint32_t size)
{
    for (uint32_t p = 0; p < size; p++)
        for (uint32_t j = 0; j < c->k; j++) {
            uint64_t q = c->q[j];
            size_t off = ((size_t)p * c->k + j) * c->n;
            for (uint32_t i = 0; i < c->n; i++) a[off + i] = addmod(a[off + i], b[off + i], q);
        }
}
/* Evaluator::sub; fhe_image.h:208. */
void orc_sub(const orc_ctx *c, uint64_t *a, const uint64_t *b, uint32_t size)
{
    for (uint32_t p = 0; p < size; p++)
        for (uint32_t j = 0; j < c->k; j++) {
            uint64_t q = c->q[j];
            size_t off = ((size_t)p * c->k + j) * c->n;
            for (uint32_t i = 0; i < c->n; i++) a[off + i] = submod(a[off + i], b[off + i], q);
        }
}
/* Evaluator::negate; fhe_resize.h:196. */
void orc_negate(const orc_ctx *c, uint64_t *a, uint32_t size)
{
    for (uint32_t p = 0; p < size; p++)
        for (uint32_t j = 0; j < c->k; j++) {
            uint64_t q = c->q[j];
            size_t off = ((size_t)p * c->k + j) * c->n;
            for (uint32_t i = 0; i < c->n; i++) a[off + i] = a[off + i] ? q - a[off + i] : 0;
        }
}

/* Encryptor::preencrypt restated: Delta*m_i (+ q mod t when m_i is in the
 * upper half) per limb, added to (sign=+1) or subtracted from (sign=-1) poly 0. */
static void plain_scale_accumulate(const orc_ctx *c, uint64_t *a, const uint64_t *plain,
                                   uint32_t plain_len, int sign)
{
    for (uint32_t j = 0; j < c->k; j++) {
        uint64_t q = c->q[j];
        uint64_t *dst = a + (size_t)j * c->n; /* poly 0, limb j */
        for (uint32_t i = 0; i < plain_len && i < c->n; i++) {
            uint64_t m = plain[i];
            if (!m) continue;
            u128 z = (u128)c->delta[j] * m;
            if (m >= c->plain_upper_half_threshold) z += c->q_mod_t;
            uint64_t s = barrett128((uint64_t)z, (uint64_t)(z >> 64), q, c->ratio0[j], c->ratio1[j]);
            dst[i] = sign > 0 ? addmod(dst[i], s, q) : submod(dst[i], s, q);
        }
    }
}
/* Evaluator::add_plain; fhe_resize.h:196, fhe_decode.h:57. */
void orc_add_plain(const orc_ctx *c, uint64_t *a, const uint64_t *plain, uint32_t plain_len)
{
    plain_scale_accumulate(c, a, plain, plain_len, +1);
}
/* Evaluator::sub_plain; fhe_image.h:317. */
void orc_sub_plain(const orc_ctx *c, uint64_t *a, const uint64_t *plain, uint32_t plain_len)
{
    plain_scale_accumulate(c, a, plain, plain_len, -1);
}

/* Evaluator::multiply_plain; fhe_image.h:221 and 41 more call sites.
 * Centred lift m -> m + (q_j - t) for m >= (t+1)/2; monomial shortcut (negacyclic shift x
 * scalar) when plain has exactly one non-zero coefficient, else NTT(plain) once per limb and
 * per poly/limb NTT -> dyadic product -> inverse NTT. */
int orc_multiply_plain(const orc_ctx *c, uint64_t *a, uint32_t size, const uint64_t *plain,
                       uint32_t plain_len)
{
    uint32_t n = c->n, nz = 0, mono = 0;
    if (plain_len > n) plain_len = n;
    for (uint32_t i = 0; i < plain_len; i++)
        if (plain[i]) {
            nz++;
            mono = i;
        }
    if (!nz) return -1;
    if (nz == 1) {
        uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * n);
        for (uint32_t j = 0; j < c->k; j++) {
            uint64_t q = c->q[j], m = plain[mono];
            if (m >= c->plain_upper_half_threshold) m += q - c->t;
            for (uint32_t p = 0; p < size; p++) {
                uint64_t *poly = a + ((size_t)p * c->k + j) * n;
                /* negacyclic_multiply_poly_mono_coeffmod: x^i * x^mono, sign flip past x^n */
                for (uint32_t i = 0; i < n; i++) {
                    uint64_t v = mulmod(c, j, poly[i], m);
                    uint32_t d = i + mono;
                    if (d >= n) {
                        d -= n;
                        v = v ? q - v : 0;
                    }
                    tmp[d] = v;
                }
                memcpy(poly, tmp, sizeof(uint64_t) * n);
            }
        }
        free(tmp);
        return 0;
    }
    uint64_t *lift = (uint64_t *)malloc(sizeof(uint64_t) * n);
    for (uint32_t j = 0; j < c->k; j++) {
        uint64_t q = c->q[j];
        for (uint32_t i = 0; i < n; i++) {
            uint64_t m = i < plain_len ? plain[i] : 0;
            lift[i] = m >= c->plain_upper_half_threshold ? m + (q - c->t) : m;
        }
        orc_ntt_forward(c, j, lift);
        for (uint32_t p = 0; p < size; p++) {
            uint64_t *poly = a + ((size_t)p * c->k + j) * n;
            orc_ntt_forward(c, j, poly);
            for (uint32_t i = 0; i < n; i++) poly[i] = mulmod(c, j, poly[i], lift[i]);
            orc_ntt_inverse(c, j, poly);
        }
    }
    free(lift);
    return 0;
}

/* ------------------------------------------------------------------ encoders
 * FractionalEncoder(t, poly_modulus, int_coeffs, frac_coeffs, base=2), server_jpeg.cpp:100,
 * restated from the SEAL 2.3 manual: integer part in binary at x^0.., i-th
 * fractional binary digit d_i (weight 2^-(i+1)) stored as -d_i at x^(n-1-i); negative
 * values have every digit negated; all coefficients mod t. */

void orc_int_encode(const orc_ctx *c, int64_t value, uint64_t *plain_out)
{
    memset(plain_out, 0, sizeof(uint64_t) * c->n);
    uint64_t mag = value < 0 ? (uint64_t)(-(value + 1)) + 1 : (uint64_t)value;
    uint64_t one = value < 0 ? c->t - 1 : 1;
    for (uint32_t i = 0; mag; i++, mag >>= 1)
        if (mag & 1) plain_out[i] = one;
}

void orc_frac_encode(const orc_ctx *c, double value, uint32_t int_coeffs, uint32_t frac_coeffs,
                     uint64_t *plain_out)
{
    (void)int_coeffs;
    int64_t vi = (int64_t)value; /* truncation toward zero */
    orc_int_encode(c, vi, plain_out);
    value -= (double)vi;
    if (value == 0) return;
    for (uint32_t i = 0; i < frac_coeffs; i++) {
        value *= 2.0;
        int64_t d = (int64_t)value;
        value -= (double)d;
        if (d > 0) plain_out[c->n - 1 - i] = c->t - 1; /* -(+1) */
        else if (d < 0) plain_out[c->n - 1 - i] = 1;   /* -(-1) */
    }
}

double orc_frac_decode(const orc_ctx *c, const uint64_t *plain)
{
    uint32_t n = c->n;
    double acc = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t m = plain[i];
        if (!m) continue;
        double v = m >= c->plain_upper_half_threshold ? -(double)(c->t - m) : (double)m;
        if (i < n / 2) acc += ldexp(v, (int)i);
        else acc -= ldexp(v, -(int)(n - i));
    }
    return acc;
}

/* ------------------------------------------------------------------ reference op graphs */

typedef struct {
    const orc_ctx *c;
    uint32_t ic, fc;
    size_t words;
    uint64_t *plain;
} gctx;

static inline void ct_copy(const gctx *g, uint64_t *dst, const uint64_t *src)
{
    memcpy(dst, src, g->words * sizeof(uint64_t));
}
/* `Ciphertext boazN(a); evaluator.add(boazN, b); dst = boazN;` (fhe_image.h:207) */
static void g_add(const gctx *g, uint64_t *dst, const uint64_t *a, const uint64_t *b, uint64_t *boaz)
{
    ct_copy(g, boaz, a);
    orc_add(g->c, boaz, b, 2);
    ct_copy(g, dst, boaz);
}
static void g_sub(const gctx *g, uint64_t *dst, const uint64_t *a, const uint64_t *b, uint64_t *boaz)
{
    ct_copy(g, boaz, a);
    orc_sub(g->c, boaz, b, 2);
    ct_copy(g, dst, boaz);
}
/* `Ciphertext boazN(a); evaluator.multiply_plain(boazN, encoder.encode(v)); dst = boazN;` */
static void g_mulp(const gctx *g, uint64_t *dst, const uint64_t *a, double v, uint64_t *boaz)
{
    ct_copy(g, boaz, a);
    orc_frac_encode(g->c, v, g->ic, g->fc, g->plain);
    orc_multiply_plain(g->c, boaz, 2, g->plain, g->c->n);
    ct_copy(g, dst, boaz);
}

/* One 1-D pass of the LLM/jpeg-6b "islow" network on d[0..7] (pointers to 8 ciphertexts).
 * scale8 = 0: row pass fhe_image.h:206-244; scale8 = 1: column pass :246-284 (each output
 * additionally multiplied by encode(0.125)). */
static void dct_pass(const gctx *g, uint64_t **d, int scale8, uint64_t *scratch)
{
    size_t W = g->words;
    uint64_t *tmp0 = scratch + 0 * W, *tmp1 = scratch + 1 * W, *tmp2 = scratch + 2 * W,
             *tmp3 = scratch + 3 * W, *tmp4 = scratch + 4 * W, *tmp5 = scratch + 5 * W,
             *tmp6 = scratch + 6 * W, *tmp7 = scratch + 7 * W, *tmp10 = scratch + 8 * W,
             *tmp11 = scratch + 9 * W, *tmp12 = scratch + 10 * W, *tmp13 = scratch + 11 * W,
             *z1 = scratch + 12 * W, *z2 = scratch + 13 * W, *z3 = scratch + 14 * W,
             *z4 = scratch + 15 * W, *z5 = scratch + 16 * W, *bz = scratch + 17 * W,
             *u = scratch + 18 * W;
    g_add(g, tmp0, d[0], d[7], bz);
    g_sub(g, tmp7, d[0], d[7], bz);
    g_add(g, tmp1, d[1], d[6], bz);
    g_sub(g, tmp6, d[1], d[6], bz);
    g_add(g, tmp2, d[2], d[5], bz);
    g_sub(g, tmp5, d[2], d[5], bz);
    g_add(g, tmp3, d[3], d[4], bz);
    g_sub(g, tmp4, d[3], d[4], bz);
    g_add(g, tmp10, tmp0, tmp3, bz);
    g_sub(g, tmp13, tmp0, tmp3, bz);
    g_add(g, tmp11, tmp1, tmp2, bz);
    g_sub(g, tmp12, tmp1, tmp2, bz);
    if (!scale8) {
        g_add(g, d[0], tmp10, tmp11, bz);
        g_sub(g, d[4], tmp10, tmp11, bz);
    } else {
        g_add(g, u, tmp10, tmp11, bz);
        g_mulp(g, d[0], u, 0.125, bz);
        g_sub(g, u, tmp10, tmp11, bz);
        g_mulp(g, d[4], u, 0.125, bz);
    }
    g_add(g, u, tmp12, tmp13, bz);
    g_mulp(g, z1, u, 0.541196100, bz);
    g_mulp(g, u, tmp13, 0.765366865, bz);
    if (!scale8) g_add(g, d[2], z1, u, bz);
    else {
        g_add(g, u, z1, u, bz);
        g_mulp(g, d[2], u, 0.125, bz);
    }
    g_mulp(g, u, tmp12, -1.847759065, bz);
    if (!scale8) g_add(g, d[6], z1, u, bz);
    else {
        g_add(g, u, z1, u, bz);
        g_mulp(g, d[6], u, 0.125, bz);
    }
    g_add(g, z1, tmp4, tmp7, bz);
    g_add(g, z2, tmp5, tmp6, bz);
    g_add(g, z3, tmp4, tmp6, bz);
    g_add(g, z4, tmp5, tmp7, bz);
    g_add(g, u, z3, z4, bz);
    g_mulp(g, z5, u, 1.175875602, bz);
    g_mulp(g, tmp4, tmp4, 0.298631336, bz);
    g_mulp(g, tmp5, tmp5, 2.053119869, bz);
    g_mulp(g, tmp6, tmp6, 3.072711026, bz);
    g_mulp(g, tmp7, tmp7, 1.501321110, bz);
    g_mulp(g, z1, z1, -0.899976223, bz);
    g_mulp(g, z2, z2, -2.562915447, bz);
    g_mulp(g, z3, z3, -1.961570560, bz);
    g_mulp(g, z4, z4, -0.390180644, bz);
    g_add(g, z3, z3, z5, bz);
    g_add(g, z4, z4, z5, bz);
    struct {
        uint64_t *t, *za, *zb;
        int out;
    } odd[4] = {{tmp4, z1, z3, 7}, {tmp5, z2, z4, 5}, {tmp6, z2, z3, 3}, {tmp7, z1, z4, 1}};
    for (int o = 0; o < 4; o++) {
        g_add(g, u, odd[o].t, odd[o].za, bz);
        if (!scale8) g_add(g, d[odd[o].out], u, odd[o].zb, bz);
        else {
            g_add(g, u, u, odd[o].zb, bz);
            g_mulp(g, d[odd[o].out], u, 0.125, bz);
        }
    }
}

/* encrypted_dct, fhe_image.h:196-288. data = 64 size-2 ciphertexts, row-major. */
void orc_encrypted_dct(const orc_ctx *c, uint64_t *data, uint32_t int_coeffs, uint32_t frac_coeffs)
{
    gctx g = {c, int_coeffs, frac_coeffs, ct_words(c, 2), NULL};
    g.plain = (uint64_t *)malloc(sizeof(uint64_t) * c->n);
    uint64_t *scratch = (uint64_t *)malloc(sizeof(uint64_t) * g.words * 19);
    uint64_t *d[8];
    for (int r = 0; r < 8; r++) {
        for (int i = 0; i < 8; i++) d[i] = data + (size_t)(8 * r + i) * g.words;
        dct_pass(&g, d, 0, scratch);
    }
    for (int col = 0; col < 8; col++) {
        for (int i = 0; i < 8; i++) d[i] = data + (size_t)(8 * i + col) * g.words;
        dct_pass(&g, d, 1, scratch);
    }
    free(scratch);
    free(g.plain);
}

/* quantize_fhe, fhe_image.h:294-305: data[i] *= encode(1/quant[i]). */
void orc_quantize_fhe(const orc_ctx *c, uint64_t *data, const double *quant, uint32_t int_coeffs,
                      uint32_t frac_coeffs)
{
    size_t W = ct_words(c, 2);
    uint64_t *plain = (uint64_t *)malloc(sizeof(uint64_t) * c->n);
    for (int i = 0; i < 64; i++) {
        orc_frac_encode(c, 1 / quant[i], int_coeffs, frac_coeffs, plain);
        orc_multiply_plain(c, data + (size_t)i * W, 2, plain, c->n);
    }
    free(plain);
}

/* rgb_to_ycc_fhe, fhe_image.h:310-325. r,g,b size-2 ciphertexts, replaced by Y,U,V. */
void orc_rgb_to_ycc_fhe(const orc_ctx *c, uint64_t *r, uint64_t *gg, uint64_t *b,
                        uint32_t int_coeffs, uint32_t frac_coeffs)
{
    gctx g = {c, int_coeffs, frac_coeffs, ct_words(c, 2), NULL};
    size_t W = g.words;
    g.plain = (uint64_t *)malloc(sizeof(uint64_t) * c->n);
    uint64_t *s = (uint64_t *)malloc(sizeof(uint64_t) * W * 6);
    uint64_t *a = s, *bb = s + W, *bz = s + 2 * W, *y = s + 3 * W, *u = s + 4 * W, *v = s + 5 * W;
    /* :317  y = r*.299 + g*.587 + b*.114 - 128 */
    g_mulp(&g, a, r, 0.299, bz);
    g_mulp(&g, bb, gg, 0.587, bz);
    g_add(&g, a, a, bb, bz);
    g_mulp(&g, bb, b, 0.114, bz);
    g_add(&g, a, a, bb, bz);
    ct_copy(&g, bz, a);
    orc_frac_encode(c, 128.0, int_coeffs, frac_coeffs, g.plain);
    orc_sub_plain(c, bz, g.plain, c->n);
    ct_copy(&g, y, bz);
    /* :318  u = r*(-.168736) - g*.331264 + b*.5 */
    g_mulp(&g, a, r, -0.168736, bz);
    g_mulp(&g, bb, gg, 0.331264, bz);
    g_sub(&g, a, a, bb, bz);
    g_mulp(&g, bb, b, 0.5, bz);
    g_add(&g, u, a, bb, bz);
    /* :319  v = r*.5 - g*.418688 - b*.081312 */
    g_mulp(&g, a, r, 0.5, bz);
    g_mulp(&g, bb, gg, 0.418688, bz);
    g_sub(&g, a, a, bb, bz);
    g_mulp(&g, bb, b, 0.081312, bz);
    g_sub(&g, v, a, bb, bz);
    ct_copy(&g, r, y);
    ct_copy(&g, gg, u);
    ct_copy(&g, b, v);
    free(s);
    free(g.plain);
}

/* ------------------------------------------------------------------ synthetic inputs */

static inline uint64_t splitmix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_fill_uniform(const orc_ctx *c, uint64_t *dst, uint64_t first_ct, uint64_t n_ct,
                      uint32_t size, uint64_t seed)
{
    uint64_t per_ct = (uint64_t)size * c->k * c->n;
    for (uint64_t ct = 0; ct < n_ct; ct++)
        for (uint32_t p = 0; p < size; p++)
            for (uint32_t j = 0; j < c->k; j++) {
                uint64_t base = ((first_ct + ct) * size + p) * c->k + j;
                uint64_t *out = dst + ct * per_ct + ((uint64_t)p * c->k + j) * c->n;
                for (uint32_t i = 0; i < c->n; i++) {
                    uint64_t g = base * c->n + i;
                    uint64_t r = splitmix64(seed + g * 0x9E3779B97F4A7C15ull);
                    out[i] = mulhi64(r, c->q[j]);
                }
            }
}

uint64_t orc_checksum(const uint64_t *v, uint64_t count)
{
    uint64_t acc = 0;
    for (uint64_t i = 0; i < count; i++) acc += splitmix64(v[i] + i * 0x9E3779B97F4A7C15ull);
    return acc;
}

/* ------------------------------------------------------------------ CPU baseline */

typedef struct {
    const orc_ctx *c;
    uint64_t *data;
    uint64_t nblocks;
    const double *quant;
    uint32_t ic, fc;
    uint64_t next; /* atomically claimed block counter */
} bench_job;

static void *bench_worker(void *arg)
{
    bench_job *jb = (bench_job *)arg;
    size_t blk = ct_words(jb->c, 2) * 64;
    for (;;) {
        uint64_t b = __atomic_fetch_add(&jb->next, 1, __ATOMIC_RELAXED);
        if (b >= jb->nblocks) break;
        orc_encrypted_dct(jb->c, jb->data + (size_t)b * blk, jb->ic, jb->fc);
        orc_quantize_fhe(jb->c, jb->data + (size_t)b * blk, jb->quant, jb->ic, jb->fc);
    }
    return NULL;
}

double orc_bench_dct_quant(const orc_ctx *c, uint64_t *data, uint64_t nblocks, const double *quant,
                           uint32_t int_coeffs, uint32_t frac_coeffs, int nthreads)
{
    bench_job jb = {c, data, nblocks, quant, int_coeffs, frac_coeffs, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, bench_worker, &jb);
    bench_worker(&jb);
    for (int i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
