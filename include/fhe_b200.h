/*
 * fhe_b200.h -- C-ABI of libfhe_b200.so: the B200-native BFV ciphertext-arithmetic hot path
 * that homo/server_jpeg.cpp (and, next, server_resize.cpp / server_decode.cpp) of
 * wfus/Fully-Homomorphic-Image-Processing drive through seal::Evaluator.
 *
 * The reference has no FFI layer: its boundary is the C++ calling convention between the
 * image-op headers (homo/fhe_image.h, fhe_resize.h, fhe_decode.h) and seal::Evaluator /
 * seal::FractionalEncoder.  Each entry point below names the reference
 * interface it replaces (file:line under /root/reference).  INTEGRATION.md shows the binding a
 * maintainer adds on the reference side (host/seal_shim.hpp gives seal-shaped C++ classes over
 * this ABI so fhe_image.h-style graphs compile unchanged in spirit).
 *
 * Conventions: extern "C", opaque handles, plain pointers and sizes; int status return
 * (0 = FHE_OK, negative = error, text via fhe_last_error(), thread-local); no exceptions cross
 * the ABI; the caller owns every buffer it passes; one CUDA stream per context, calls are
 * asynchronous with respect to the host unless stated (fhe_ctx_sync waits); a context is
 * thread-compatible (one thread at a time), there is no global mutable state.
 *
 * Ciphertext layout (host and device): [ciphertext][poly][limb][n] uint64, residues in
 * [0, q_limb).  A size-2 ciphertext at n=4096, k=2 is 131072 bytes.  (SEAL 2.3 stores n+1
 * coefficients per limb with a zero top coefficient; drop it when importing, Appendix A.1.)
 * Plaintext layout: n uint64 coefficients in [0, t).
 *
 * There is NO CPU fallback: every compute entry point fails with FHE_ERR_CUDA when no CUDA
 * device is usable.
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHE_OK 0
#define FHE_ERR_INVALID_ARGUMENT (-1) /* seal: std::invalid_argument */
#define FHE_ERR_CUDA (-2)
#define FHE_ERR_UNSUPPORTED (-3)
#define FHE_ERR_ZERO_PLAIN (-4) /* seal: multiply_plain throws on an all-zero plaintext */
#define FHE_ERR_ALLOC (-5)

typedef struct fhe_ctx fhe_ctx;           /* seal::SEALContext + seal::Evaluator */
typedef struct fhe_ptab fhe_ptab;         /* a set of pre-transformed seal::Plaintext */
typedef struct fhe_dct_plan fhe_dct_plan; /* tables of encrypted_dct + quantize_fhe */

const char *fhe_last_error(void);
int fhe_version(void);

/* ---- context -------------------------------------------------------------------------------
 * Replaces EncryptionParameters / SEALContext / Evaluator construction,
 * homo/server_jpeg.cpp:74-80,93 (set_poly_modulus "1x^n + 1", set_coeff_modulus, set_plain_modulus).
 * q: k distinct primes < 2^61, each 1 mod 2n.  cuda_stream: a cudaStream_t to enqueue on, or
 * NULL to let the context create its own. */
int fhe_ctx_create(uint32_t n, uint32_t k, const uint64_t *q, uint64_t t, int device,
                   void *cuda_stream, fhe_ctx **out);
void fhe_ctx_destroy(fhe_ctx *ctx);
int fhe_ctx_sync(fhe_ctx *ctx);
/* seal::coeff_modulus_128(n), server_jpeg.cpp:78 ([SEAL-recall] chain). */
int fhe_default_coeff_modulus(uint32_t n, uint64_t *q_out, uint32_t *k_out);
/* bytes of one ciphertext of `size` polys */
size_t fhe_ct_bytes(const fhe_ctx *ctx, uint32_t size);

/* ---- buffers ------------------------------------------------------------------------------ */
int fhe_device_alloc(fhe_ctx *ctx, size_t bytes, void **dptr);
int fhe_device_free(fhe_ctx *ctx, void *dptr);
int fhe_host_alloc(size_t bytes, void **hptr); /* pinned */
int fhe_host_free(void *hptr);
int fhe_upload(fhe_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);   /* synchronous */
int fhe_download(fhe_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes); /* synchronous */
/* device -> device on the context's stream: seal::Ciphertext copy construction / assignment
 * (`Ciphertext boaz1(data[0])`, fhe_image.h:207). */
int fhe_device_copy(fhe_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);
/* `rows` runs of `width` bytes, `*_pitch` bytes apart (device -> device, on the stream): gathers
 * ciphertexts into a batch, or re-strides a batch of size-s ciphertexts into size-s' storage (what
 * Evaluator::add does when the sizes differ). */
int fhe_device_copy_2d(fhe_ctx *ctx, void *dst_dev, size_t dst_pitch, const void *src_dev, size_t src_pitch,
                       size_t width, size_t rows);

/* ---- encoder (host logic) -----------------------------------------------------------------
 * seal::FractionalEncoder(plain_modulus, poly_modulus, int_coeffs, frac_coeffs, base).encode,
 * constructed at server_jpeg.cpp:100 and called at 43 sites of fhe_image.h (e.g. :221).
 * base must be 2 (the reference's --base override is dead code, server_jpeg.cpp:66).
 * Pure host logic (n = ring degree, t = plain modulus): needs no device. */
int fhe_frac_encode(uint32_t n, uint64_t t, double value, uint32_t int_coeffs, uint32_t frac_coeffs,
                    uint32_t base, uint64_t *plain_out /* n */);

/* ---- Evaluator: exact ring ops on device arrays of n_ct ciphertexts -----------------------
 * Evaluator::add / sub / negate, e.g. fhe_image.h:207-208, fhe_resize.h:196.  In place on a. */
int fhe_add(fhe_ctx *ctx, uint64_t *a, const uint64_t *b, uint64_t n_ct, uint32_t size);
int fhe_sub(fhe_ctx *ctx, uint64_t *a, const uint64_t *b, uint64_t n_ct, uint32_t size);
int fhe_negate(fhe_ctx *ctx, uint64_t *a, uint64_t n_ct, uint32_t size);

/* Plaintext tables: for each of `count` plaintexts (host, [count][n], coefficients < t) the
 * centred lift is transformed once and kept in Shoup form, with its Delta-scaled polynomial
 * (Encryptor::preencrypt) beside it.  Replaces the per-call work inside
 * Evaluator::multiply_plain / add_plain / sub_plain. */
int fhe_ptab_create(fhe_ctx *ctx, const uint64_t *plain_coeffs, uint32_t count, fhe_ptab **out);
void fhe_ptab_destroy(fhe_ptab *tab);
/* Evaluator::multiply_plain(ct, plain), fhe_image.h:221 and 41 more: every ciphertext of the
 * array times plaintext `index` (one fused NTT -> dyadic -> inverse-NTT kernel per limb-poly).
 * FHE_ERR_ZERO_PLAIN if that plaintext is zero. */
int fhe_multiply_plain(fhe_ctx *ctx, uint64_t *a, uint64_t n_ct, uint32_t size, const fhe_ptab *tab,
                       uint32_t index);
/* Evaluator::add_plain / sub_plain (fhe_resize.h:196 / fhe_image.h:317): poly 0 +/- Delta*m. */
int fhe_add_plain(fhe_ctx *ctx, uint64_t *a, uint64_t n_ct, uint32_t size, const fhe_ptab *tab,
                  uint32_t index);
int fhe_sub_plain(fhe_ctx *ctx, uint64_t *a, uint64_t n_ct, uint32_t size, const fhe_ptab *tab,
                  uint32_t index);

/* util/smallntt ntt_negacyclic_harvey / inverse ([SEAL-recall]) over n_limbpolys consecutive
 * [limb-cycled] polynomials, in place.  The NTT-domain coefficient order is internal to this
 * library (a fixed permutation of SEAL's bit-reversed order, see csrc/ntt.cuh). */
int fhe_ntt_forward(fhe_ctx *ctx, uint64_t *a, uint64_t n_limbpolys);
/* That permutation, for callers that want to look at NTT-domain data (the parity tests do): SEAL's
 * scrambled index j is stored at position (j mod E) * (n / E) + j div E with E = 2^fhe_ntt_loge(n),
 * the number of coefficients a thread keeps in registers.  0 for an unsupported n. */
uint32_t fhe_ntt_loge(uint32_t n);
int fhe_ntt_inverse(fhe_ctx *ctx, uint64_t *a, uint64_t n_limbpolys);

/* ---- the north-star path: encrypted 8x8 DCT + quantize ------------------------------------
 * Plan = the 12 distinct constants of encrypted_dct in the order of fhe_image.h:221-236
 * (0.541196100, 0.765366865, -1.847759065, 1.175875602, 0.298631336, 2.053119869, 3.072711026,
 * 1.501321110, -0.899976223, -2.562915447, -1.961570560, -0.390180644) as plaintexts
 * [12][n], encode(0.125) (:259) [n], and optionally the 64 plaintexts encode(1/quant[i])
 * of quantize_fhe (:301) [64][n]; quant == NULL gives encrypted_dct alone. */
int fhe_dct_plan_create(fhe_ctx *ctx, const uint64_t *dct_consts, const uint64_t *eighth,
                        const uint64_t *quant, fhe_dct_plan **out);
void fhe_dct_plan_destroy(fhe_dct_plan *plan);
/* encrypted_dct(data, ...) [+ quantize_fhe(data, quant, ...)] on nblocks blocks, each 64
 * size-2 ciphertexts row-major (fhe_image.h:192-199, 294-298; block loop server_jpeg.cpp:113-138).
 * d_in, d_out: device, [block][64][2][k][n]; may alias (in place, as in the reference). */
int fhe_dct8x8_quantize(fhe_ctx *ctx, const fhe_dct_plan *plan, const uint64_t *d_in,
                        uint64_t *d_out, uint64_t nblocks);
/* Same with HOST buffers (what server_jpeg.cpp holds after Ciphertext::load, :115-124):
 * chunks are copied in, transformed and copied back on three overlapped streams.
 * Synchronous: returns when h_out is complete.  Pinned buffers (fhe_host_alloc) run faster. */
int fhe_dct8x8_quantize_host(fhe_ctx *ctx, const fhe_dct_plan *plan, const uint64_t *h_in,
                             uint64_t *h_out, uint64_t nblocks);

/* ---- the whole server_jpeg tile step -------------------------------------------------------
 * server_jpeg.cpp:113-138 per 8x8 RGB tile: rgb_to_ycc_fhe on the 64 pixels (fhe_image.h:310-325:
 * Y = .299 R + .587 G + .114 B - 128, Cb = -.168736 R - .331264 G + .5 B,
 * Cr = .5 R - .418688 G - .081312 B) then encrypted_dct of Y, Cb, Cr (:196-288) [then quantize_fhe
 * (:294-305) when tables are given], fused: one forward NTT kernel with the colour conversion in
 * its epilogue, the pointwise DCT per channel, one inverse NTT kernel.
 * color_consts [9][n]: encode() of 0.299, 0.587, 0.114, -0.168736, 0.331264, 0.5, 0.5, 0.418688,
 * 0.081312 in this order (the reference subtracts the terms it encodes positively, :318-319);
 * offset128 [n] = encode(128.0); dct_consts / eighth as in fhe_dct_plan_create; quant_y / quant_c:
 * [64][n] encode(1/quant[i]) for Y and for Cb/Cr, or NULL (server_jpeg.cpp itself never quantises).
 * Tiles: [tile][3][64][2][k][n] -- R64, G64, B64 in (server_jpeg.cpp:115-124), Y64, Cb64, Cr64 out
 * (:146-153); d_in may equal d_out. */
typedef struct fhe_jpeg_plan fhe_jpeg_plan;
int fhe_jpeg_plan_create(fhe_ctx *ctx, const uint64_t *color_consts, const uint64_t *offset128,
                         const uint64_t *dct_consts, const uint64_t *eighth, const uint64_t *quant_y,
                         const uint64_t *quant_c, fhe_jpeg_plan **out);
void fhe_jpeg_plan_destroy(fhe_jpeg_plan *plan);
int fhe_jpeg_rgb_tiles(fhe_ctx *ctx, const fhe_jpeg_plan *plan, const uint64_t *d_in, uint64_t *d_out,
                       uint64_t ntiles);

/* All-zero ciphertext storage on the context's stream: the transparent encryption of 0 that the
 * deterministic server-side Encryptor starts from (fhe_resize.h:230-266 and fhe_decode.h:54,134
 * encrypt public scalars on the server). */
int fhe_device_zero(fhe_ctx *ctx, void *dptr, size_t bytes);
/* The transparent (key-less, noise-free) encryption of `count` plaintexts in one call:
 * d_dst[i] = (Delta * m_i, 0), size 2, plaintext i at plain[i * n .. ) with coefficients < t.
 * Same ciphertext as fhe_device_zero + fhe_add_plain, without building a plaintext table (one
 * upload, one kernel, no device synchronisation): the per-pixel fractions of fhe_resize.h:230-266
 * are encrypted this way on the server. */
int fhe_encrypt_transparent(fhe_ctx *ctx, const uint64_t *plain, uint32_t count, uint64_t *d_dst);

/* ---- Evaluator: ciphertext x ciphertext ----------------------------------------------------
 * Evaluator::multiply(a, b) (fhe_resize.h:175-179,197-198; fhe_decode.h:71-97,235,239): BFV tensor
 * product with the BEHZ full-RNS scaling round(t/q .) (SEAL 2.3 [SEAL-recall], SURVEY.md Appendix
 * A.6).  The auxiliary base is this library's own (55-bit primes, csrc/behz.cuh): the result of the
 * procedure does not depend on it, and is compared bit for bit with the oracle's SEAL base.  n_ct independent pairs:
 * a [n_ct][size_a][k][n], b [n_ct][size_b][k][n] -> dest [n_ct][size_a + size_b - 1][k][n].
 * The reference's in-place form `evaluator.multiply(a, b)` is dest -> a afterwards (sizes differ,
 * so dest must be a separate buffer and must not overlap the operands).  Sizes up to 32
 * (the decode graph reaches 21 x 2, fhe_decode.h:239).
 * FHE_ERR_UNSUPPORTED for more than 8 coeff_modulus primes. */
int fhe_multiply(fhe_ctx *ctx, const uint64_t *a, uint32_t size_a, const uint64_t *b, uint32_t size_b,
                 uint64_t *dest, uint64_t n_ct);
/* Evaluator::square(a) (fhe_resize.h:174, fhe_decode.h:67...): same result as multiply(a, a), the
 * operand is extended and transformed once. */
int fhe_square(fhe_ctx *ctx, const uint64_t *a, uint32_t size_a, uint64_t *dest, uint64_t n_ct);

/* Operand forms.  Evaluator::multiply transforms each operand (BEHZ extension q -> B_sk, forward NTT in
 * every modulus of q and B_sk) before the tensor step.  The reference's graphs multiply by the same
 * ciphertext over and over -- t, t2 and "t3" of Cubic are shared by the 4 rows x 3 channels of an output
 * pixel (fhe_resize.h:174-179, 296-303), shifted_x by four products of each Taylor series
 * (fhe_decode.h:79-97) -- so the transformed operand is exposed as a handle: fhe_form_create does the
 * operand half of fhe_multiply once for n_ct ciphertexts [n_ct][size][k][n]; fhe_multiply_form does the
 * rest for n_ct products, product i taking ciphertext i / a_div of `a` and i / b_div of `b` (a_div = 1:
 * one operand per product; 3: three consecutive products share one).  dest
 * [n_ct][a.size + b.size - 1][k][n] is bit-identical to fhe_multiply on the same ciphertexts (same
 * arithmetic, transformed once).  a == b (same handle, same divisor) is Evaluator::square. */
typedef struct fhe_form fhe_form;
int fhe_form_create(fhe_ctx *ctx, const uint64_t *d_ct, uint32_t size, uint64_t n_ct, fhe_form **out);
void fhe_form_destroy(fhe_form *form);
int fhe_multiply_form(fhe_ctx *ctx, const fhe_form *a, uint32_t a_div, const fhe_form *b, uint32_t b_div,
                      uint64_t *dest, uint64_t n_ct);

/* Asynchronous host <-> device transfers (pinned host memory: fhe_host_alloc) on the context's copy
 * streams, for pipelines that stream rows or chunks through the device: a transfer starts after the work
 * queued on the context's stream at the time of the call and overlaps what is queued afterwards.
 * fhe_transfer_fence: (uploads) the context's stream waits for every upload issued so far -- call it
 * before the first kernel that reads an uploaded buffer; (downloads) the HOST waits until all downloads
 * except the keep_downloads most recent ones have landed -- call it before the source of an older
 * download is freed or overwritten, and with 0 before reading the last results on the host.
 * fhe_upload_fence: the HOST waits until all uploads except the keep_uploads most recent ones have
 * completed -- call it before the pinned source of an older upload is refilled (a reader thread that
 * parses the next chunk of a ciphertext stream into a ring of pinned buffers, server/stream_b200.hpp). */
int fhe_upload_async(fhe_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int fhe_download_async(fhe_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int fhe_transfer_fence(fhe_ctx *ctx, uint32_t keep_downloads);
int fhe_upload_fence(fhe_ctx *ctx, uint32_t keep_uploads);

/* count copies of one ciphertext (size polynomials) into consecutive elements of d_dst */
int fhe_device_broadcast(fhe_ctx *ctx, uint64_t *d_dst, const uint64_t *d_src, uint32_t size, uint64_t count);
/* Evaluator::add_plain with a DIFFERENT short plaintext per ciphertext: element i of a ([n_ct][size][k][n])
 * += plaintext i, given by its first ncoef coefficients plains[i][0 .. ncoef) (host; the rest are zero).
 * The step approximation adds encode(i) to its running offset once per (position, j) (fhe_decode.h:229):
 * the offsets of all positions are the first one plus short integer plaintexts, added in one launch. */
int fhe_add_plain_each(fhe_ctx *ctx, uint64_t *a, uint64_t n_ct, uint32_t size, const uint64_t *plains, uint32_t ncoef);

/* ---- resize graphs: Cubic and Linear (fhe_resize.h:143-189, :191-204) as fused device graphs --------
 * The same ciphertexts as the reference's op-by-op evaluation (every step is exact arithmetic with
 * canonical results and BEHZ multiply is a deterministic function of its operands), computed for n_ct
 * independent calls at once in 6 (Cubic) / 5 (Linear) kernel launches:
 *   gather + linear combinations + BEHZ extension -> forward transforms -> tensor products with the
 *   CACHED forms of the fraction powers fused with the inverse transforms -> BEHZ scaling + sum
 *   (+ x encode(0.5) + d).
 * taps: HOST array [n_ct][4] (Cubic: A B C D) or [n_ct][2] (Linear: A B) of DEVICE pointers to
 * ciphertexts of size_in polynomials ([size_in][k][n]); the 4 x 4 neighbourhood of an output pixel
 * (fhe_resize.h:269-290) is read in place through the table, never copied.
 * Cubic: t (size 2), t2 = square(t) and t3 = multiply(t, t) (sic, fhe_resize.h:174-175; equal sizes)
 * as forms; call i uses ciphertext i / t_div of each.  dest [n_ct][size_in + t2.size - 1][k][n].
 * Linear: one_minus_t = negate(t) + encode(1.0) and t as forms; dest [n_ct][size_in + t.size - 1][k][n].
 * fhe_resize_plan_create: plains = [5][n] coefficients of encode(3), encode(2), encode(5), encode(4),
 * encode(0.5) from the caller's FractionalEncoder (fhe_resize.h:151-183); FHE_ERR_UNSUPPORTED when
 * they are not short +-1 polynomials (binary FractionalEncoder encodings always are) -- the caller
 * then composes the graph from the Evaluator entry points. */
typedef struct fhe_resize_plan fhe_resize_plan;
int fhe_resize_plan_create(fhe_ctx *ctx, const uint64_t *plains, fhe_resize_plan **out);
void fhe_resize_plan_destroy(fhe_resize_plan *plan);
int fhe_cubic(fhe_ctx *ctx, const fhe_resize_plan *plan, const uint64_t *const *taps, uint32_t size_in,
              const fhe_form *t, const fhe_form *t2, const fhe_form *t3, uint32_t t_div, uint64_t *dest,
              uint64_t n_ct);
int fhe_linear(fhe_ctx *ctx, const uint64_t *const *taps, uint32_t size_in, const fhe_form *one_minus_t,
               const fhe_form *t, uint32_t t_div, uint64_t *dest, uint64_t n_ct);

/* Evaluator::relinearize(ct, evk) (no call site in homo/; keys at tests/parameters.cpp:89-90,
 * decomposition_bit_count = DBC fhe_image.h:28).
 * fhe_evk_create takes the keys of KeyGenerator::generate_evaluation_keys(dbc, count) in
 * COEFFICIENT form (host): [count][k][digits][2][k][n], key r for s^(2+r), limb i, digit l < digits
 * = fhe_evk_digits(ctx, dbc) = max_i ceil(bits(q_i) / dbc); the pair encrypts
 * s^(2+r) 2^(l dbc) in limb i only.  (SEAL stores them NTT-transformed in its own order.) */
typedef struct fhe_evk fhe_evk;
uint32_t fhe_evk_digits(const fhe_ctx *ctx, uint32_t dbc);
int fhe_evk_create(fhe_ctx *ctx, const uint64_t *evk_host, uint32_t dbc, uint32_t count, fhe_evk **out);
void fhe_evk_destroy(fhe_evk *evk);
/* a [n_ct][size][k][n] -> dest [n_ct][2][k][n]; needs count >= size - 2; dest must not alias a
 * unless size == 2. */
int fhe_relinearize(fhe_ctx *ctx, const uint64_t *a, uint32_t size, const fhe_evk *evk, uint64_t *dest,
                    uint64_t n_ct);

/* ---- synthetic inputs / checks ---------------------------------------------
 * Counter-based SplitMix64 -> uniform residues, same generator as oracle/fhe_oracle.c. */
int fhe_fill_uniform(fhe_ctx *ctx, uint64_t *d_dst, uint64_t first_ct, uint64_t n_ct, uint32_t size,
                     uint64_t seed);
/* order-sensitive 64-bit checksum of a device u64 array; synchronous. */
int fhe_checksum(fhe_ctx *ctx, const uint64_t *d, uint64_t count, uint64_t *out);

/* ---- tuning / introspection ---------------------------------------------------------------
 * Kernels this library launched on this context since creation (all streams). */
uint64_t fhe_launch_count(const fhe_ctx *ctx);
/* blocks per internal chunk of fhe_dct8x8_quantize (0 = whole call in one chunk). */
int fhe_set_dct_chunk_blocks(fhe_ctx *ctx, uint32_t blocks);
/* Per-kernel device timing of fhe_dct8x8_quantize (CUDA events recorded on the context's stream
 * between its kernels; the reference times the same region with steady_clock, fhe_image.h:200,285).
 * fhe_get_stage_times waits for the stream and returns the milliseconds of the LAST call:
 * ms[0] forward NTT, ms[1] pointwise DCT(+quantize), ms[2] inverse NTT (summed over chunks). */
int fhe_set_stage_timing(fhe_ctx *ctx, int enabled);
int fhe_get_stage_times(fhe_ctx *ctx, float *ms /* 3 */);

/* Per-kernel device time of everything this context launches between the two calls: an event is
 * recorded after every kernel launch (a launch's time = the gap to the previous mark on the context's
 * stream, so host-side gaps between launches are attributed to the kernel that follows them; use on
 * warmed-up, GPU-bound sequences, outside timed regions).  fhe_profile_end waits for the stream and
 * returns, per kernel name in order of first appearance: names as one '\n'-separated string, summed
 * milliseconds and launch counts. */
int fhe_profile_begin(fhe_ctx *ctx);
int fhe_profile_end(fhe_ctx *ctx, char *names, size_t names_len, float *ms, uint64_t *counts,
                    uint32_t max_kernels, uint32_t *n_out);

#ifdef __cplusplus
}
#endif
#endif
