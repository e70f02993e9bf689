// ubench.cu -- B200 pipe-rate probes for the modular-multiply design (profiles/r01d_ubench.txt).
// Each kernel runs NCHAIN independent dependency chains per thread for ITER iterations; the host
// reports warp-instructions per cycle per SM sub-partition for the op under test.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../fully-homomorphic-image-processing_b200/csrc/modarith.cuh"

// ---------------------------------------------------------------------------------------------
// mulcd (EXPERIMENT, not used by the library: measured no faster than mulc on B200 because DFMA and
// IMAD do not overlap -- profiles/r01d_ubench.txt): the same lazy multiplication by a precomputed constant with the QUOTIENT estimated on
// the FP64 pipe, which the integer path leaves idle (B200 issues DFMA at the IMAD rate on its
// own pipe; profiles/r01d_ubench.txt).  x = xh 2^31 + xl (xl < 2^31, xh < 2^31 for x < 2^62) and
// x w = xh wh + xl w (mod q) with wh = w 2^31 mod q, so the quotient
//     Q = round(xl (w/q) + xh (wh/q))  < 2^32
// needs two DFMA on exact 31-bit integers (errors: two roundings to integer at magnitude 2^52
// plus 2^-22 from the constants => |Q - exact| <= 1 + 2^-22), and the remainder
//     r = xl w + xh wh + 2q - Q q   (mod 2^64)  in [q(1 - 2^-22), q(3 + 2^-22)]  subset (0, 4q)
// needs 3 IMAD.WIDE + 3 IMAD instead of mulc's 5 + 4.  Requires x < 2^62, q < 2^55.
struct alignas(16) W4 {
    u64 w, wh;      // w, w * 2^31 mod q
    double al, ah;  // w / q, wh / q (round to nearest)
};
inline W4 make_w4(u64 w, u64 q)
{
    W4 c;
    c.w = w;
    c.wh = (u64)((((unsigned __int128)w) << 31) % q);
    c.al = (double)((long double)c.w / (long double)q);
    c.ah = (double)((long double)c.wh / (long double)q);
    return c;
}
FHE_HD double u32_to_double_magic(u32 v)  // 2^52 + v, exactly
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)v);
#else
    const u64 bits = 0x4330000000000000ull | v;
    double d;
    __builtin_memcpy(&d, &bits, 8);
    return d;
#endif
}
FHE_HD u64 mulcd(u64 x, W4 c, u64 q2, u64 nq)
{
    const u32 xl = (u32)x & 0x7fffffffu, xh = (u32)(x >> 31);
    const double Xl = u32_to_double_magic(xl) - 4503599627370496.0;  // exact
    const double Xh = u32_to_double_magic(xh) - 4503599627370496.0;
#if defined(__CUDA_ARCH__)
    const double V = __fma_rn(Xh, c.ah, __fma_rn(Xl, c.al, 6755399441055744.0));  // 2^52 + 2^51
    const u32 Q = (u32)__double2loint(V);
    const u32 w0 = (u32)c.w, w1 = (u32)(c.w >> 32), h0 = (u32)c.wh, h1 = (u32)(c.wh >> 32);
    const u32 n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 r;
    asm("{\n\t"
        ".reg .u32 lo, hi;\n\t"
        ".reg .u64 acc;\n\t"
        "mad.wide.u32 acc, %1, %3, %9;\n\t"   // xl*w0 + 2q
        "mad.wide.u32 acc, %2, %5, acc;\n\t"  // + xh*wh0
        "mad.wide.u32 acc, %7, %8, acc;\n\t"  // + Q*nq0
        "mov.b64 {lo, hi}, acc;\n\t"
        "mad.lo.u32 hi, %1, %4, hi;\n\t"      // + 2^32 (xl*w1 + xh*wh1 + Q*nq1)
        "mad.lo.u32 hi, %2, %6, hi;\n\t"
        "mad.lo.u32 hi, %7, %10, hi;\n\t"
        "mov.b64 %0, {lo, hi};\n\t"
        "}"
        : "=l"(r)
        : "r"(xl), "r"(xh), "r"(w0), "r"(w1), "r"(h0), "r"(h1), "r"(Q), "r"(n0), "l"(q2), "r"(n1));
    return r;
#else
    const double V = __builtin_fma(Xh, c.ah, __builtin_fma(Xl, c.al, 6755399441055744.0));
    u64 bits;
    __builtin_memcpy(&bits, &V, 8);
    const u32 Q = (u32)bits;
    return (u64)xl * c.w + (u64)xh * c.wh + q2 + (u64)Q * nq;
#endif
}


#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int NCHAIN = 8, ITER = 2048;

__global__ void k_imad_wide(u64 *out, u32 a, u32 b)
{
    u64 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(v[i]) : "r"((u32)v[i] ^ a), "r"(b));
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad(u64 *out, u32 a, u32 b)
{
    u32 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(a), "r"(b));
    u32 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(u64 *out, double a, double b)
{
    double v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i)
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(a), "d"(b));
    double s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (u64)s;
}
// one DFMA per IMAD.WIDE: do the two pipes overlap?
__global__ void k_mix(u64 *out, u32 a, u32 b, double c, double d)
{
    u64 v[NCHAIN];
    double f[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) { v[i] = threadIdx.x + i; f[i] = i; }
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(v[i]) : "r"((u32)v[i] ^ a), "r"(b));
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(c), "d"(d));
        }
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i] + (u64)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_i2f(u64 *out, u32 a)
{
    u32 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            double d;
            asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d) : "r"(v[i]));
            v[i] = (u32)__double2hiint(d) + a;
        }
    u32 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mulc(u64 *out, W2 c, u64 n2q, LimbArith L)
{
    u64 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x * 977 + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) v[i] = mulc(v[i], c, n2q);
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mulcd(u64 *out, W4 c, LimbArith L)
{
    u64 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x * 977 + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) v[i] = mulcd(v[i], c, 2 * L.q, L.nq);
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// variant with the magic-constant cancellation folded into a per-constant K (40-byte table entry):
// two DFMA, no DADD
struct W5 { u64 w, wh; double al, ah, K; };
__device__ __forceinline__ u64 mulcdk(u64 x, W5 c, u64 q2, u64 nq)
{
    const u32 xl = (u32)x & 0x7fffffffu, xh = (u32)(x >> 31);
    const double V = __fma_rn(u32_to_double_magic(xh), c.ah, __fma_rn(u32_to_double_magic(xl), c.al, c.K));
    const u32 Q = (u32)__double2loint(V);
    return (u64)xl * c.w + (u64)xh * c.wh + q2 + (u64)Q * nq;
}
__global__ void k_mulcdk(u64 *out, W5 c, LimbArith L)
{
    u64 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x * 977 + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) v[i] = mulcdk(v[i], c, 2 * L.q, L.nq);
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// variant converting with I2F
__device__ __forceinline__ u64 mulcdi(u64 x, W4 c, u64 q2, u64 nq)
{
    const u32 xl = (u32)x & 0x7fffffffu, xh = (u32)(x >> 31);
    const double V = __fma_rn(__uint2double_rn(xh), c.ah, __fma_rn(__uint2double_rn(xl), c.al, 6755399441055744.0));
    const u32 Q = (u32)__double2loint(V);
    return (u64)xl * c.w + (u64)xh * c.wh + q2 + (u64)Q * nq;
}
__global__ void k_mulcdi(u64 *out, W4 c, LimbArith L)
{
    u64 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x * 977 + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) v[i] = mulcdi(v[i], c, 2 * L.q, L.nq);
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// butterfly-shaped: mulc + add + sub with offset, as in ntt.cuh fwd_stage
__global__ void k_bfly(u64 *out, W2 c, LimbArith L)
{
    u64 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x * 977 + i;
    for (int it = 0; it < ITER / 2; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; i += 2) {
            u64 x = v[i], t = mulc(v[i + 1], c, L.n2q);
            v[i] = reduce63(x + t, L);
            v[i + 1] = x + L.q5 - t;
        }
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bflyd(u64 *out, W4 c, LimbArith L)
{
    u64 v[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) v[i] = threadIdx.x * 977 + i;
    for (int it = 0; it < ITER / 2; ++it)
#pragma unroll
        for (int i = 0; i < NCHAIN; i += 2) {
            u64 x = v[i], t = mulcd(v[i + 1], c, 2 * L.q, L.nq);
            v[i] = reduce63(x + t, L);
            v[i + 1] = x + L.q5 - t;
        }
    u64 s = 0;
    for (int i = 0; i < NCHAIN; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static void run(const char *name, double ops_per_thread, F launch)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    int clk_khz = 0, sms = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // grid = sms * 4 CTAs of 256 threads = 32 warps per SM = 8 warps per SMSP
    const double warp_ops_per_smsp = ops_per_thread * 8.0;
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-10s %8.3f ms  %7.3f cycles per warp-op per SMSP (at nominal %d MHz)  %.1f Gop/s/GPU\n", name, ms,
           cycles / warp_ops_per_smsp, clk_khz / 1000, ops_per_thread * sms * 4 * 256 / (ms * 1e-3) * 1e-9);
}

int main()
{
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 4, threads = 256;
    u64 *out;
    CK(cudaMalloc(&out, (size_t)grid * threads * 8));
    const u64 q = 0x7fffffff380001ull;
    LimbArith L;
    L.q = q; L.n2q = 0 - 2 * q; L.nq = 0 - q; L.q5 = 5 * q;
    L.mu63 = (u32)((((unsigned __int128)1) << 63) / q);
    L.mu64 = (u32)((((unsigned __int128)1) << 64) / q);
    const u64 w = 0x123456789abcdeull % q;
    W2 c2; c2.w = w; c2.p = (u64)((((unsigned __int128)w) << 63) / q);
    W4 c4 = make_w4(w, q);
    const double ops = (double)NCHAIN * ITER;
    run("imad.wide", ops, [&] { k_imad_wide<<<grid, threads>>>(out, 12345u, 0x9e3779b9u); });
    run("imad.lo", ops, [&] { k_imad<<<grid, threads>>>(out, 12345u, 0x9e3779b9u); });
    run("dfma", ops, [&] { k_dfma<<<grid, threads>>>(out, 1.0000001, 0.5); });
    run("wide+dfma", ops, [&] { k_mix<<<grid, threads>>>(out, 12345u, 0x9e3779b9u, 1.0000001, 0.5); });
    run("i2f.f64", ops, [&] { k_i2f<<<grid, threads>>>(out, 3u); });
    run("mulc", ops, [&] { k_mulc<<<grid, threads>>>(out, c2, L.n2q, L); });
    run("mulcd", ops, [&] { k_mulcd<<<grid, threads>>>(out, c4, L); });
    W5 c5; c5.w = c4.w; c5.wh = c4.wh; c5.al = c4.al; c5.ah = c4.ah;
    c5.K = 6755399441055744.0 - 4503599627370496.0 * c4.al - 4503599627370496.0 * c4.ah;
    run("mulcdk", ops, [&] { k_mulcdk<<<grid, threads>>>(out, c5, L); });
    run("mulcdi", ops, [&] { k_mulcdi<<<grid, threads>>>(out, c4, L); });
    run("bfly", ops / 2, [&] { k_bfly<<<grid, threads>>>(out, c2, L); });
    run("bflyd", ops / 2, [&] { k_bflyd<<<grid, threads>>>(out, c4, L); });
    // correctness spot check of mulcd against exact arithmetic
    {
        unsigned long long bad = 0;
        u64 x = 0x1234567ull;
        for (int i = 0; i < 2000000; ++i) {
            x = splitmix64(x) >> 2;  // < 2^62
            const u64 r = mulcd(x, c4, 2 * q, L.nq);
            const u64 ex = (u64)(((unsigned __int128)x * w) % q);
            if (r % q != ex || r >= 4 * q) ++bad;
        }
        printf("mulcd host check: %llu bad of 2000000\n", bad);
    }
    CK(cudaFree(out));
    return 0;
}
