// dct.cuh -- encrypted 8x8 DCT (+ quantize) as a pointwise linear map in the NTT domain.
//
// Replaces the op-by-op graph of encrypted_dct (homo/fhe_image.h:196-288: 8 row passes
// :206-244, 8 column passes with the extra x encode(0.125) :246-284) followed by quantize_fhe
// (:294-305).  Every op there is an exact ring operation in R_q (ct +/- ct, ct x plain), so for
// each (ciphertext poly, limb, NTT index) the whole graph is a fixed 64 -> 64 linear map over
// Z_q whose coefficients are the NTT values of the lifted plaintext constants
// (SURVEY.md 0.6, 7.3).  Canonical residues are unique, hence the result is bit-identical to
// SEAL's op-by-op evaluation.  The butterfly network below is the reference's own (the LLM
// "islow" network, mirror of dct() at fhe_image.h:400-484); only the arithmetic is lazy:
// additions never reduce, subtractions add a compile-time multiple of q, and the multiplications
// (modarith.cuh mulc) by the 12 constants / 64 merged scale factors bring values back under 5q.
// With q < 2^55 (256 q <= 2^63) the whole 2-D network needs no reduction step at all.
#pragma once
#include "modarith.cuh"

namespace fhe {

// order of the 12 DCT constants in the plan tables (fhe_image.h:221-236)
enum DctConst {
    C_0541196100 = 0, C_0765366865, C_M1847759065, C_1175875602, C_0298631336, C_2053119869,
    C_3072711026, C_1501321110, C_M0899976223, C_M2562915447, C_M1961570560, C_M0390180644,
    DCT_NCONST
};

// One 1-D pass on d[0..7].  Inputs < INB*q.  Outputs: d0,d4 < 8*INB*q; d2,d6 < 10q; odd < 20q.
// Needs 8*INB <= FHE_BMAX (every mulc argument stays below 2^63).
// c: the 12 constants, constant i at c[i * CS] (CS = 1: a thread-private array; CS = DCT_TILE: a
// shared-memory table [12][DCT_TILE] addressed from the thread's own position).
template <int INB, int CS = 1>
FHE_HD void dct8_pass(u64 *d, const W2 *c, u64 q, u64 n2q)
{
    static_assert(8 * INB <= 256, "dct8_pass: lazy bound exceeds 2^63");
    const u64 o1 = (u64)INB * q, o2 = 2 * o1, o4 = 4 * o1;
    u64 tmp0 = d[0] + d[7], tmp7 = d[0] - d[7] + o1;
    u64 tmp1 = d[1] + d[6], tmp6 = d[1] - d[6] + o1;
    u64 tmp2 = d[2] + d[5], tmp5 = d[2] - d[5] + o1;
    u64 tmp3 = d[3] + d[4], tmp4 = d[3] - d[4] + o1;
    u64 tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3 + o2;
    u64 tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2 + o2;
    d[0] = tmp10 + tmp11;
    d[4] = tmp10 - tmp11 + o4;
    u64 z1 = mulc(tmp12 + tmp13, c[C_0541196100 * CS], n2q);
    d[2] = z1 + mulc(tmp13, c[C_0765366865 * CS], n2q);
    d[6] = z1 + mulc(tmp12, c[C_M1847759065 * CS], n2q);
    z1 = tmp4 + tmp7;
    u64 z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    u64 z5 = mulc(z3 + z4, c[C_1175875602 * CS], n2q);
    tmp4 = mulc(tmp4, c[C_0298631336 * CS], n2q);
    tmp5 = mulc(tmp5, c[C_2053119869 * CS], n2q);
    tmp6 = mulc(tmp6, c[C_3072711026 * CS], n2q);
    tmp7 = mulc(tmp7, c[C_1501321110 * CS], n2q);
    z1 = mulc(z1, c[C_M0899976223 * CS], n2q);
    z2 = mulc(z2, c[C_M2562915447 * CS], n2q);
    z3 = mulc(z3, c[C_M1961570560 * CS], n2q) + z5;
    z4 = mulc(z4, c[C_M0390180644 * CS], n2q) + z5;
    d[7] = tmp4 + z1 + z3;
    d[5] = tmp5 + z2 + z4;
    d[3] = tmp6 + z2 + z3;
    d[1] = tmp7 + z1 + z4;
}

// Tile geometry of the pointwise kernel: a CTA of 8 warps works on TILE = 32 consecutive NTT
// positions of one limb; warp w owns row w in the row phase and column w in the column phase,
// lane = NTT position.  The 8x8 transposition between the phases goes through shared memory
// (64 x 32 u64 = 16 KiB per buffer, conflict-free: a warp always touches 32 consecutive words).
constexpr int DCT_TILE = 32;
// bound (units of q) of the forward-NTT output the row phase reads (NTT_OUT_REDUCED: [0, 4q))
constexpr int DCT_ROW_INB = 4;

// Row phase for (row r, lane): block rows are ciphertexts 8r..8r+7 (fhe_image.h:192-195).
// in: pointer to (ct 0, this poly, this limb, first position of the tile); ct_stride in u64.
FHE_HD void dct_row_load(int r, int lane, const u64 *__restrict__ in, size_t ct_stride, u64 *d)
{
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = in[(size_t)(8 * r + i) * ct_stride + lane];
}
template <int CS = 1>
FHE_HD void dct_row_compute(int r, int lane, u64 *d, u64 *sbuf, const W2 *c, u64 q, u64 n2q)
{
    dct8_pass<DCT_ROW_INB, CS>(d, c, q, n2q);
#pragma unroll
    for (int i = 0; i < 8; ++i) sbuf[(8 * r + i) * DCT_TILE + lane] = d[i];
}
template <int CS = 1>
FHE_HD void dct_row_phase(int r, int lane, const u64 *__restrict__ in, size_t ct_stride, u64 *sbuf,
                          const W2 *c, u64 q, u64 n2q)
{
    u64 d[8];
    dct_row_load(r, lane, in, ct_stride, d);
    dct_row_compute<CS>(r, lane, d, sbuf, c, q, n2q);
}

// Column phase for (column col, lane): 1-D pass down the column, then one multiplication per
// output by the merged scale constant  s[i] = NTT(0.125) * NTT(1/quant[8i+col])  (or NTT(0.125)
// alone when no quantisation table is attached), times n^-1 so the inverse NTT that follows skips
// its own scaling.  Output < 5q (inverse-NTT input range).
// s: scale constants of outputs (row i, this column), s[i * s_stride]
template <int CS = 1>
FHE_HD void dct_col_phase(int col, int lane, u64 *__restrict__ out, size_t ct_stride, const u64 *sbuf,
                          const W2 *c, const W2 *__restrict__ s, size_t s_stride, u64 q, u64 n2q)
{
    u64 d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = sbuf[(8 * i + col) * DCT_TILE + lane];
    dct8_pass<8 * DCT_ROW_INB, CS>(d, c, q, n2q);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(size_t)(8 * i + col) * ct_stride + lane] = mulc(d[i], s[i * s_stride], n2q);
}

}  // namespace fhe
