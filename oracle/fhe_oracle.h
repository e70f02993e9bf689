/*
 * fhe_oracle.h -- CPU oracle for the BFV ciphertext-arithmetic hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (fully-homomorphic-image-processing_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED.  The arithmetic the reference drives lives in the git
 * submodule SEAL (https://github.com/wfus/SEAL @ 467638eb60b2..., "version
 * 2.3", /root/reference/.gitmodules:1-3, README.md:70) which is EMPTY in
 * /root/reference, and the reference's tests hold no ciphertext-level golden
 * vectors (tests are print-only programs).  This file therefore restates the
 * published SEAL 2.3 / BFV algorithm (Fan-Vercauteren 2012; Bajard et al.
 * 2016 full-RNS variant; Harvey 2014 NTT butterflies; SEAL 2.3.0 manual) and
 * is anchored on the reference's own call sites:
 *   homo/fhe_image.h:196-288  encrypted_dct      -> orc_encrypted_dct
 *   homo/fhe_image.h:294-305  quantize_fhe       -> orc_quantize_fhe
 *   homo/fhe_image.h:310-325  rgb_to_ycc_fhe     -> orc_rgb_to_ycc_fhe
 *   homo/server_jpeg.cpp:74-80,100  parameters   -> orc_ctx_create / encoder
 * It is cross-checked against an independent pure-Python big-int model
 * (oracle/pymodel.py) and algebraic identities (Dec o Eval == Eval_plain), and
 * -- at the level of decrypted results, the only level at which the reference
 * pins anything -- against every number the reference holds for this path:
 *   tests/parameters.cpp:100-128   42x^3 - 27x + 1 at 12345 = 79017370138936
 *   benchmark/results.txt:37-42    JPEG pipeline RMSError 1.71767 (boazbarak.jpg)
 *   benchmark/results.txt:17-21    bilinear resize RMSError 17.9597
 * (tests/test_reference_known_answers.py, test_client_jpeg.py, test_resize_published.py
 * and their GPU counterparts).  Ciphertext BITS remain unpinned: SEAL is absent.
 *
 * Ciphertext layout everywhere: [poly][limb][n] u64, residues in [0, q_limb).
 * (SEAL 2.3 stores n+1 coefficients per limb with a zero top coefficient; the
 *  padding coefficient is dropped here.)
 */
#ifndef FHE_ORACLE_H
#define FHE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LIMBS 16

typedef struct orc_ctx orc_ctx;

/* n = ring degree (power of two), k primes q[j] == 1 (mod 2n), plain modulus t.
 * Mirrors EncryptionParameters/SEALContext set up at server_jpeg.cpp:74-80. */
orc_ctx *orc_ctx_create(uint32_t n, uint32_t k, const uint64_t *q, uint64_t t);
void orc_ctx_destroy(orc_ctx *c);

uint32_t orc_n(const orc_ctx *c);
uint32_t orc_k(const orc_ctx *c);
uint64_t orc_q(const orc_ctx *c, uint32_t j);
uint64_t orc_t(const orc_ctx *c);
uint64_t orc_psi(const orc_ctx *c, uint32_t j);       /* minimal primitive 2n-th root */
uint64_t orc_delta_mod(const orc_ctx *c, uint32_t j); /* floor(q/t) mod q_j            */
uint64_t orc_q_mod_t(const orc_ctx *c);               /* "upper_half_increment"        */

/* Default SEAL 2.3.1 coeff_modulus_128(n) chain; returns k. */
uint32_t orc_default_coeff_modulus(uint32_t n, uint64_t *q_out);

/* Negacyclic NTT per limb, in place. forward: natural -> bit-reversed order. */
void orc_ntt_forward(const orc_ctx *c, uint32_t j, uint64_t *a);
void orc_ntt_inverse(const orc_ctx *c, uint32_t j, uint64_t *a);
/* out = a * b in Z_qj[x]/(x^n+1); out may alias a or b. */
void orc_negacyclic_mul(const orc_ctx *c, uint32_t j, const uint64_t *a, const uint64_t *b,
                        uint64_t *out);

/* ---- Evaluator subset (SEAL 2.3 semantics) ----
 * All in place on the first argument, like seal::Evaluator. */
void orc_add(const orc_ctx *c, uint64_t *a, const uint64_t *b, uint32_t size);
void orc_sub(const orc_ctx *c, uint64_t *a, const uint64_t *b, uint32_t size);
void orc_negate(const orc_ctx *c, uint64_t *a, uint32_t size);
/* plain: plain_len coefficients in [0,t). */
void orc_add_plain(const orc_ctx *c, uint64_t *a, const uint64_t *plain, uint32_t plain_len);
void orc_sub_plain(const orc_ctx *c, uint64_t *a, const uint64_t *plain, uint32_t plain_len);
/* returns 0, or -1 if plain is zero (SEAL throws invalid_argument). */
int orc_multiply_plain(const orc_ctx *c, uint64_t *a, uint32_t size, const uint64_t *plain,
                       uint32_t plain_len);

/* ---- FractionalEncoder(t, poly_modulus, int_coeffs, frac_coeffs, base) ----
 * server_jpeg.cpp:100. plain_out has n coefficients. base 2 only (binary). */
void orc_frac_encode(const orc_ctx *c, double value, uint32_t int_coeffs, uint32_t frac_coeffs,
                     uint64_t *plain_out);
void orc_int_encode(const orc_ctx *c, int64_t value, uint64_t *plain_out);
/* Test-harness decode: centred coefficients, x^i -> 2^i for i < n/2 and
 * x^(n-i) -> -2^-i otherwise. */
double orc_frac_decode(const orc_ctx *c, const uint64_t *plain);

/* ---- Reference op graphs, op by op with the reference's copies ----
 * data: 64 size-2 ciphertexts, row-major 8x8 (fhe_image.h:192-195). */
void orc_encrypted_dct(const orc_ctx *c, uint64_t *data, uint32_t int_coeffs, uint32_t frac_coeffs);
void orc_quantize_fhe(const orc_ctx *c, uint64_t *data, const double *quant, uint32_t int_coeffs,
                      uint32_t frac_coeffs);
void orc_rgb_to_ycc_fhe(const orc_ctx *c, uint64_t *r, uint64_t *g, uint64_t *b,
                        uint32_t int_coeffs, uint32_t frac_coeffs);

/* ---- Synthetic inputs: counter-based SplitMix64 -> uniform residues.
 * Fills n_ct ciphertexts of `size` polys starting at global ciphertext index first_ct. */
void orc_fill_uniform(const orc_ctx *c, uint64_t *dst, uint64_t first_ct, uint64_t n_ct,
                      uint32_t size, uint64_t seed);
/* 64-bit order-sensitive checksum of a u64 array (sum of mix(i, v)). */
uint64_t orc_checksum(const uint64_t *v, uint64_t count);

/* ---- CPU baseline: DCT+quantize on nblocks blocks (64 cts each), op by op,
 * blocks spread over nthreads pthreads. Returns seconds. */
double orc_bench_dct_quant(const orc_ctx *c, uint64_t *data, uint64_t nblocks, const double *quant,
                           uint32_t int_coeffs, uint32_t frac_coeffs, int nthreads);

/* ==== part 2 (fhe_oracle_behz.c): ciphertext x ciphertext ops and the graphs built on them ====
 * BEHZ full-RNS multiply/square and relinearize restated from SURVEY.md Appendix A.6-A.7
 * ([SEAL-recall]; PARITY UNPINNED), anchored on fhe_resize.h:174-179,197-198 and
 * fhe_decode.h:67-97,147-177,235,239. */
typedef struct orc_behz orc_behz;
orc_behz *orc_behz_create(const orc_ctx *c);
void orc_behz_destroy(orc_behz *b);
uint32_t orc_behz_aux_count(const orc_behz *b);           /* |B| */
uint64_t orc_behz_bsk(const orc_behz *b, uint32_t j);      /* B_sk primes, m_sk last */
/* steps 1-2 (fastbconv_mtilde + mont_rq): x [k][n] in base q -> out [|B|+1][n] in base Bsk */
void orc_behz_extend(const orc_behz *b, const uint64_t *x, uint64_t *out);
/* steps 6-8 (times t, fast_floor, fastbconv_sk): xq [k][n], xb [|B|+1][n] -> out [k][n] */
void orc_behz_scale_round(const orc_behz *b, const uint64_t *xq, const uint64_t *xb, uint64_t *out);
/* Evaluator::multiply / square; dest has s1 + s2 - 1 (2 s - 1) polys, must not alias inputs. */
void orc_multiply(const orc_behz *b, const uint64_t *a1, uint32_t s1, const uint64_t *a2, uint32_t s2,
                  uint64_t *dest);
void orc_square(const orc_behz *b, const uint64_t *a, uint32_t s, uint64_t *dest);

/* harness (client side in the reference): keys, encryption, decryption phase, evaluation keys.
 * sk [k][n]; pk [2][k][n]; all polynomials in coefficient form. */
void orc_keygen(const orc_behz *b, uint64_t seed, uint64_t *sk, uint64_t *pk);
void orc_encrypt(const orc_behz *b, const uint64_t *pk, const uint64_t *plain, uint32_t plain_len,
                 uint64_t seed, uint64_t *ct);
void orc_phase(const orc_behz *b, const uint64_t *sk, const uint64_t *ct, uint32_t size, uint64_t *out);
uint32_t orc_evk_digits(const orc_behz *b, uint32_t dbc);
/* evk layout [count][k][digits][2][k][n], coefficient form */
void orc_evk_gen(const orc_behz *b, const uint64_t *sk, uint32_t dbc, uint32_t count, uint64_t seed,
                 uint64_t *evk);
void orc_relinearize(const orc_behz *b, const uint64_t *ct, uint32_t size, const uint64_t *evk,
                     uint32_t dbc, uint32_t evk_count, uint64_t *dest);

/* reference graphs; each returns the size of the result ciphertext(s) */
uint32_t orc_cubic(const orc_behz *b, const uint64_t *const in[4], const uint32_t sz[4],
                   const uint64_t *t, uint32_t st, uint32_t ic, uint32_t fc, uint64_t *out);
uint32_t orc_linear(const orc_behz *b, const uint64_t *A, uint32_t sA, const uint64_t *B, uint32_t sB,
                    const uint64_t *t, uint32_t st, uint32_t ic, uint32_t fc, uint64_t *out);
uint32_t orc_homomorphic_sincos(const orc_behz *b, const uint64_t *x, uint32_t sx, const uint64_t *zero,
                                int is_cos, uint32_t ic, uint32_t fc, uint64_t *out);
uint32_t orc_approximated_step(const orc_behz *b, const uint64_t *amplitude, const uint64_t *index,
                               const uint64_t *count, int order, int degree, double delta, int npos,
                               const uint64_t *zeros, uint32_t ic, uint32_t fc, uint64_t *run_out);

#ifdef __cplusplus
}
#endif
#endif
