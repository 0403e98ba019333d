/*
 * tc_spike.cu -- FEASIBILITY SPIKE, not product code (VERDICT r1 "Next" #5): can Blackwell's 5th-generation tensor
 * cores (tcgen05.mma kind::i8, accumulators in TMEM) carry the first DFT-32 pass of the spectrum chain / a FIR stage?
 *
 * What it measures on one B200:
 *   (1) exactness: D[128 x N] (s32, TMEM) = A[128 x 32] (UNSIGNED 8-bit: raw I/Q bytes need no conversion) x
 *       B[N x 32]^T (signed 8-bit: one 8-bit slice of a DFT-32 / tap matrix), read back with tcgen05.ld and compared
 *       with the integer product computed on the host -- bit for bit;
 *   (2) rate: back-to-back accumulating MMAs on shared-memory resident operands, one issuing thread per SM, all SMs:
 *       dense int8 TOP/s for N = 64 and N = 256;
 *   (3) the DFT-32 numerics: cos/sin matrices cut into 3 signed 8-bit slices (2^-7, 2^-14, 2^-21), three exact integer
 *       products recombined in fp32 vs a float64 DFT of the same bytes.
 * DESIGN.md section 5.0b has the conclusion (measured no-go for the product path, with the reasons).
 *
 * Operand layout: K-major, no swizzle, 8 x 16-byte core matrices: element (r, k) of a [rows x 32] 8-bit tile lives at
 * (r / 8) * 256 + (k / 16) * 128 + (r % 8) * 16 + (k % 16)  ->  descriptor LBO = 128 B, SBO = 256 B.
 * Every wait is bounded (clock64): a protocol mistake makes the kernel report an error, it cannot hang the GPU.
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/tc_spike tools/tc_spike.cu
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets (all >> 4),
 * version 1 (Blackwell), no swizzle */
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46; /* version_ = 1 */
    return d;
}
/* instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at [4,6), a_format [7,10) 0 = unsigned 8-bit,
 * b_format [10,13) 1 = signed 8-bit, both K-major, N >> 3 at [17,23), M >> 4 at [24,29) */
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t a_fmt, uint32_t b_fmt)
{
    return (2u << 4) | (a_fmt << 7) | (b_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* bounded mbarrier wait: false on timeout */
__device__ __forceinline__ bool bar_wait(uint64_t *bar, uint32_t parity, long long timeout)
{
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (!ok && clock64() - t0 > timeout) return false;
    }
    return true;
}

/* element (r, k) of a K-major 8-bit tile with 32 columns, canonical no-swizzle layout */
__host__ __device__ inline uint32_t tile_off(uint32_t r, uint32_t k) { return (r >> 3) * 256u + (k >> 4) * 128u + (r & 7u) * 16u + (k & 15u); }

template <int N>
__global__ void __launch_bounds__(128, 1) k_spike(const uint8_t *a_rows /* [128][32] u8 */, const int8_t *b_rows /* [N][32] s8 */,
                                                   int32_t *d_out /* [gridDim.x][128][N] or null */, uint32_t iters,
                                                   long long *cycles /* [gridDim.x] */, uint32_t *err)
{
    __shared__ __align__(1024) uint8_t s_a[128 * 32];
    __shared__ __align__(1024) uint8_t s_b[N * 32];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = (int)threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 128 * 32; i += 128) s_a[tile_off(i >> 5, i & 31)] = a_rows[i];
    for (int i = tid; i < N * 32; i += 128) s_b[tile_off(i >> 5, i & 31)] = (uint8_t)b_rows[i];
    if (warp == 0) { /* TMEM: N columns of 128 lanes x 32 bits (power of two >= 32) */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> visible to the MMA's async proxy */
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t da = make_desc(smem_u32(s_a), 128, 256), db = make_desc(smem_u32(s_b), 128, 256);
        constexpr uint32_t idesc = make_idesc(128, N, 0 /* A unsigned */, 1 /* B signed */);
        t0 = clock64();
        for (uint32_t i = 0; i < iters; ++i) mma_i8(tmem, da, db, idesc, i > 0 ? 1u : 0u);
        mma_commit(&s_bar);
    }
    const bool ok = bar_wait(&s_bar, 0, 4000000000ll);
    if (tid == 0) { t1 = clock64(); cycles[blockIdx.x] = t1 - t0; }
    if (!ok) { if (tid == 0) atomicAdd(err, 1u); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (ok && d_out) { /* warp w reads TMEM lanes 32 w .. 32 w + 31 = rows of D, 32 columns at a time */
        int32_t *row = d_out + ((size_t)blockIdx.x * 128 + tid) * N;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(addr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) row[c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)N) : "memory");
}

template <int N>
static int run(const char *label, const std::vector<uint8_t> &a, const std::vector<int8_t> &b, int sms, bool check, uint32_t iters,
               std::vector<int32_t> *d_host)
{
    uint8_t *d_a; int8_t *d_b; int32_t *d_d = nullptr; long long *d_cyc; uint32_t *d_err;
    CK(cudaMalloc(&d_a, a.size())); CK(cudaMalloc(&d_b, b.size()));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms)); CK(cudaMalloc(&d_err, 4)); CK(cudaMemset(d_err, 0, 4));
    CK(cudaMemcpy(d_a, a.data(), a.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice));
    if (check) CK(cudaMalloc(&d_d, sizeof(int32_t) * (size_t)sms * 128 * N));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_spike<N><<<sms, 128>>>(d_a, d_b, d_d, iters, d_cyc, d_err); /* warm-up */
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_spike<N><<<sms, 128>>>(d_a, d_b, d_d, iters, d_cyc, d_err);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    uint32_t err = 0; CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    std::vector<long long> cyc(sms); CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long cmax = 0; for (long long c : cyc) cmax = c > cmax ? c : cmax;
    const double ops = 2.0 * 128 * N * 32 * (double)iters * sms;
    printf("%-28s N=%3d iters=%7u  timeouts=%u  kernel %.3f ms  issue->commit %lld cycles  %.1f dense int8 TOP/s (event time), %.1f MAC/clk/SM\n",
           label, N, iters, err, ms, cmax, ops / (ms * 1e-3) / 1e12, 128.0 * N * 32 * iters / (double)cmax);
    int bad = 0;
    if (check && !err) {
        std::vector<int32_t> d((size_t)sms * 128 * N);
        CK(cudaMemcpy(d.data(), d_d, d.size() * 4, cudaMemcpyDeviceToHost));
        for (int s = 0; s < sms && bad < 5; s += (sms > 1 ? sms - 1 : 1))
            for (int r = 0; r < 128 && bad < 5; ++r)
                for (int n = 0; n < N; ++n) {
                    long long ref = 0;
                    for (int k = 0; k < 32; ++k) ref += (long long)a[r * 32 + k] * (long long)b[n * 32 + k];
                    ref *= iters;
                    if ((long long)d[((size_t)s * 128 + r) * N + n] != ref) {
                        if (bad < 5) printf("  MISMATCH sm %d row %d col %d: got %d want %lld\n", s, r, n, d[((size_t)s * 128 + r) * N + n], ref);
                        ++bad;
                    }
                }
        printf("  exactness (u8 x s8 -> s32, %d rows x %d cols x 2 SMs checked): %s\n", 128, N, bad ? "FAILED" : "bit-exact");
        if (d_host) d_host->assign(d.begin(), d.begin() + 128 * N);
    }
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_cyc); cudaFree(d_err); if (d_d) cudaFree(d_d);
    return err || bad;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sm_%d%d, %d SMs, %.0f MHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate / 1e3);
    const int sms = prop.multiProcessorCount;
    /* A: 128 "rows" of 32 raw unsigned bytes (4 frames x 32 lanes of one I/Q plane: row m = lane t of frame m / 32, K = j) */
    std::vector<uint8_t> a(128 * 32);
    uint32_t lcg = 12345u;
    for (auto &v : a) { lcg = lcg * 1664525u + 1013904223u; v = (uint8_t)(lcg >> 24); }
    a[0] = 255; a[1] = 0; a[2] = 128; a[3] = 127;
    int rc = 0;
    /* (1) + (3): the DFT-32 cos / sin matrix, three signed 8-bit slices; N = 64 columns = (k1, cos | sin) */
    const double kPi = 3.14159265358979323846;
    std::vector<int8_t> slice[3];
    std::vector<double> g(64 * 32);
    for (int n = 0; n < 64; ++n)
        for (int j = 0; j < 32; ++j) {
            const int k1 = n & 31;
            const double ang = -2.0 * kPi * (double)((j * k1) & 31) / 32.0;
            g[n * 32 + j] = n < 32 ? cos(ang) : sin(ang);
        }
    for (int s = 0; s < 3; ++s) slice[s].resize(64 * 32);
    for (int i = 0; i < 64 * 32; ++i) { /* g = q0 2^-7 + q1 2^-14 + q2 2^-21 + r, |q| <= 127 (|g| <= 1: q0 in [-127, 127] after clamp) */
        double r = g[i];
        for (int s = 0; s < 3; ++s) {
            double q = nearbyint(ldexp(r, 7 * (s + 1)));
            if (q > 127) q = 127;
            if (q < -127) q = -127;
            slice[s][i] = (int8_t)q;
            r -= ldexp(q, -7 * (s + 1));
        }
    }
    std::vector<int32_t> d[3];
    for (int s = 0; s < 3; ++s) {
        char label[64]; snprintf(label, sizeof label, "DFT-32 slice %d (2^-%d)", s, 7 * (s + 1));
        rc |= run<64>(label, a, slice[s], sms, true, 1, &d[s]);
    }
    if (!rc) { /* recombine in fp32 like an epilogue would, compare with the float64 DFT of the same bytes (offset -127.5 removed by
                  the DC column on the CPU side of the comparison: sum_j g[n][j] is 32 or 0) */
        double worst = 0, scale = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 64; ++n) {
                const float y = fmaf((float)d[0][r * 64 + n], 0x1p-7f, fmaf((float)d[1][r * 64 + n], 0x1p-14f, (float)d[2][r * 64 + n] * 0x1p-21f));
                double ref = 0;
                for (int j = 0; j < 32; ++j) ref += g[n * 32 + j] * (double)a[r * 32 + j];
                worst = fmax(worst, fabs((double)y - ref));
                scale = fmax(scale, fabs(ref));
            }
        printf("  DFT-32 pass from three exact int8 products, recombined in fp32: max abs err %.3e on values up to %.1f (%.2e relative to the largest)\n",
               worst, scale, worst / scale);
    }
    /* (2) rate */
    std::vector<int8_t> b256(256 * 32);
    for (auto &v : b256) { lcg = lcg * 1664525u + 1013904223u; v = (int8_t)(lcg >> 24); }
    rc |= run<64>("rate, 128 x 64 x 32 tiles", a, slice[0], sms, false, 4096, nullptr);
    rc |= run<256>("rate, 128 x 256 x 32 tiles", a, b256, sms, true, 16, nullptr);
    rc |= run<256>("rate, 128 x 256 x 32 tiles", a, b256, sms, false, 4096, nullptr);
    printf(rc ? "SPIKE FAILED\n" : "SPIKE OK\n");
    return rc;
}
