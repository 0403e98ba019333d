// ubench.cu -- instruction-throughput microbenchmarks on sm_100a for the design questions of
// FFMA vs FFMA2 (packed fp32x2) rate, what shares an issue slot with them, u8->f32
// conversion cost.  Prints warp-instructions per clock per SM and fp32 lane-ops per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
typedef unsigned long long u64;
#define ITERS 2048
#define REP8(x) x x x x x x x x

__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, long long *cyc, float seed)
{
    float a[8]; u64 p[8];
    float x = seed + threadIdx.x * 1e-9f, y = 0.999f;
    u64 px = pk(x, x + 1e-9f), py = pk(y, y);
    uint32_t w = threadIdx.x * 2654435761u, acc32 = 0;
    uint32_t iv[8];
    for (int i = 0; i < 8; ++i) iv[i] = w + i;
    const u64 pden = pk(__uint_as_float(0x00110000u + (threadIdx.x << 16 & 0x7f0000u)), __uint_as_float(0x00350000u)), pbig = pk(1e30f, -1e30f);
    const uint32_t hv = 0x3c003c00u + (threadIdx.x & 3); /* two f16 values ~1 */
    __shared__ u64 sm[1024];
    sm[threadIdx.x] = px; sm[threadIdx.x + 512] = py;
    for (int i = 0; i < 8; ++i) { a[i] = i * 0.1f; p[i] = pk(i * 0.1f, i * 0.2f); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) { // FFMA 3-reg
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(y));
        } else if (MODE == 1) { // FFMA2
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
        } else if (MODE == 2) { // FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(px));
        } else if (MODE == 3) { // FFMA2 + PRMT 1:1
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                asm volatile("prmt.b32 %0, %1, %2, 0x7504;" : "=r"(acc32) : "r"(w + i), "r"(acc32));
            }
        } else if (MODE == 4) { // PRMT only
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("prmt.b32 %0, %1, %0, 0x7504;" : "+r"(acc32) : "r"(w + i));
        } else if (MODE == 5) { // I2F.U8 (cvt.rn.f32.u8 of a byte)
#pragma unroll
            for (int i = 0; i < 8; ++i) { float f; asm volatile("{ .reg .b32 t; .reg .u8 b; shr.b32 t, %1, 8; cvt.u8.u32 b, t; cvt.rn.f32.u8 %0, b; }" : "=f"(f) : "r"(w + i)); a[i] += f; }
        } else if (MODE == 6) { // FFMA2 + LDS.64 1:1
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                u64 v = *((volatile u64 *)&sm[(threadIdx.x + i * 32 + it) & 1023]);
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(v), "l"(py));
            }
        } else if (MODE == 7) { // FFMA (scalar) + PRMT 1:1
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(y));
                asm volatile("prmt.b32 %0, %1, %2, 0x7504;" : "=r"(acc32) : "r"(w + i), "r"(acc32));
            }
        } else if (MODE == 8) { // FFMA2 : FFMA 1:1
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(y));
            }
        } else if (MODE == 9) { // FFMA2 with 3 distinct dependent-free operands per instr (register bandwidth)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(p[(i + 3) & 7]), "l"(p[(i + 5) & 7]));
        } else if (MODE == 10) { // LDS.64 only
#pragma unroll
            for (int i = 0; i < 8; ++i) { u64 v = *((volatile u64 *)&sm[(threadIdx.x + i * 32 + it) & 1023]); p[i] ^= v; }
        } else if (MODE == 11) { // MUFU.RCP (loop-carried through an FADD so it cannot be folded)
#pragma unroll
            for (int i = 0; i < 8; ++i) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r + 1.0f; }
        } else if (MODE == 12) { // IADD3 x8 independent chains
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(iv[i]) : "r"(w));
        } else if (MODE == 13) { // FFMA2 + IADD 1:1, all independent
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(iv[i]) : "r"(w));
            }
        } else if (MODE == 14) { // FFMA + IADD 1:1
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(y));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(iv[i]) : "r"(w));
            }
        } else if (MODE == 15) { // 2 FFMA2 + 1 IADD
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                if (i & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(iv[i]) : "r"(w));
            }
        } else if (MODE == 16) { // 2 FFMA + 1 IADD
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(y));
                if (i & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(iv[i]) : "r"(w));
            }
        } else if (MODE == 17) { // FFMA2 with scalar .F32 multiplier from a register (FIR form)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("{ .reg .b64 t; mov.b64 t, {%2, %2}; fma.rn.f32x2 %0, %1, t, %0; }" : "+l"(p[i]) : "l"(px), "f"(a[i & 3]));
        } else if (MODE == 18) { // PRMT x8 independent
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("prmt.b32 %0, %0, %1, 0x7504;" : "+r"(iv[i]) : "r"(w));
        } else if (MODE == 19) { // FFMA2 + PRMT 1:1 independent
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                asm volatile("prmt.b32 %0, %0, %1, 0x7504;" : "+r"(iv[i]) : "r"(w));
            }
        } else if (MODE == 21) { // I2F.U8 with loop-variant input: IADD + I2F.U8 + FADD per op
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float f;
                asm volatile("add.u32 %0, %0, %1;" : "+r"(iv[i]) : "r"(w));
                asm volatile("{ .reg .u8 b; cvt.u8.u32 b, %1; cvt.rn.f32.u8 %0, b; }" : "=f"(f) : "r"(iv[i]));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(f));
            }
        } else if (MODE == 22) { // same without the conversion (IADD + FADD only), the baseline to subtract
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(iv[i]) : "r"(w));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x));
            }
        } else if (MODE == 23) { // I2F.U8 + FFMA2 1:1 (does the conversion unit run beside the FMA pipe?)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float f;
                asm volatile("{ .reg .u8 b; cvt.u8.u32 b, %1; cvt.rn.f32.u8 %0, b; }" : "=f"(f) : "r"(iv[i]));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                iv[i] = __float_as_uint(f) + iv[i];
            }
        } else if (MODE == 24) { // FHADD x8 independent (f32 = f16 half of a register + f32), sm_100 mixed-precision add
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, hi, %0; }" : "+f"(a[i]) : "r"(hv));
        } else if (MODE == 25) { // FFMA2 + FHADD 1:1
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, hi, %0; }" : "+f"(a[i]) : "r"(hv));
            }
        } else if (MODE == 26) { // conversion form A next to the FIR: 2 PRMT + FADD2 + 8 FFMA2
            asm volatile("prmt.b32 %0, %0, %1, 0x7504;" : "+r"(iv[0]) : "r"(w));
            asm volatile("prmt.b32 %0, %0, %1, 0x7514;" : "+r"(iv[1]) : "r"(w));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[0]) : "l"(px));
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
        } else if (MODE == 27) { // conversion form B next to the FIR: PRMT + 2 FHADD + 8 FFMA2
            asm volatile("prmt.b32 %0, %0, %1, 0x4140;" : "+r"(iv[0]) : "r"(w));
            asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, lo, %0; }" : "+f"(a[0]) : "r"(hv));
            asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, hi, %0; }" : "+f"(a[1]) : "r"(hv));
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
        } else if (MODE == 30) { // FFMA2 whose multiplicand is a pair of SUBNORMAL floats (u8 raw form, cplx2.cuh form C)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(pden), "l"(pbig));
        } else if (MODE == 28) { // HFMA2 (f16x2) x8 independent: is the half pipe the FP32 pipe?
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(iv[i]) : "r"(hv), "r"(w));
        } else if (MODE == 29) { // FFMA2 + HFMA2 1:1
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(iv[i]) : "r"(hv), "r"(w));
            }
        } else if (MODE == 20) { // FFMA2 + LDS.32 conflict-free 2:1
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
                if (i & 1) { uint32_t v = *((volatile uint32_t *)&sm[0] + ((threadIdx.x + i * 32 + it) & 2047)); iv[i] ^= v; }
            }
        }
    }
    long long t1 = clock64();
    float s = 0; u64 q = 0;
    for (int i = 0; i < 8; ++i) { s += a[i]; q ^= p[i]; }
    uint32_t z = 0; for (int i = 0; i < 8; ++i) z ^= iv[i];
    if (s == 1234.5f || q == 77 || acc32 == 99 || z == 12345) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, double instr_per_iter, double lane_flops_per_instr)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, 512>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, 512>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += (double)h[i]; c /= 148;
    double warp_instr = 16.0 * ITERS * instr_per_iter; // 16 warps per SM
    printf("%-34s cycles %9.0f  warp-instr/clk/SM %6.3f  fp32-lane-ops/clk/SM %7.1f  (%.3f ms, %.0f MHz)\n", name, c,
           warp_instr / c, warp_instr * 32 * lane_flops_per_instr / c, ms, c / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA", 8, 1);
    run<1>("FFMA2", 8, 2);
    run<9>("FFMA2 (3 distinct pair operands)", 8, 2);
    run<2>("FADD2", 8, 2);
    run<8>("FFMA2 + FFMA 1:1", 8, 1.5);
    run<4>("PRMT", 8, 0);
    run<3>("FFMA2 + PRMT 1:1", 16, 1);
    run<7>("FFMA + PRMT 1:1", 16, 0.5);
    run<10>("LDS.64", 8, 0);
    run<6>("FFMA2 + LDS.64 1:1", 16, 1);
    run<11>("MUFU.RCP (+FADD)", 16, 0);
    run<12>("IADD x8 indep", 8, 0);
    run<18>("PRMT x8 indep", 8, 0);
    run<13>("FFMA2 + IADD 1:1 indep", 16, 1);
    run<19>("FFMA2 + PRMT 1:1 indep", 16, 1);
    run<14>("FFMA + IADD 1:1 indep", 16, 0.5);
    run<15>("FFMA2 + IADD 2:1", 12, 4.0 / 3);
    run<16>("FFMA + IADD 2:1", 12, 2.0 / 3);
    run<17>("FFMA2 scalar-bcast multiplier", 8, 2);
    run<20>("FFMA2 + LDS.32 2:1", 12, 4.0 / 3);
    run<21>("IADD + I2F.U8 + FADD (x8)", 24, 0);
    run<22>("IADD + FADD (x8) baseline", 16, 0);
    run<23>("I2F.U8 + FFMA2 + IADD (x8)", 24, 0);
    run<24>("FHADD x8 indep", 8, 1);
    run<25>("FFMA2 + FHADD 1:1", 16, 1.5);
    run<26>("cvt A: 2 PRMT + FADD2 + 8 FFMA2", 11, 18.0 / 11);
    run<27>("cvt B: PRMT + 2 FHADD + 8 FFMA2", 11, 18.0 / 11);
    run<30>("FFMA2 subnormal multiplicand", 8, 2);
    run<28>("HFMA2 x8 indep", 8, 0);
    run<29>("FFMA2 + HFMA2 1:1", 16, 1);
    return 0;
}
