/*
 * cuda_emu.h -- a tiny host-side emulation of the CUDA execution model, used ONLY by the CPU
 * test-suite (tests/test_emulated_kernels.py) to run the product's kernel source
 * (stm32f7-rtlsdr_b200/csrc/*.cuh) on the host and check its index arithmetic against the
 * oracle before GPU time is spent.  It is test infrastructure, never part of the product path:
 * the product library is built by nvcc for sm_100a and has no CPU fallback.
 *
 * Model: each CUDA thread of a block runs as a ucontext fiber; __syncthreads() yields to the
 * next fiber, so by the time a fiber resumes every other fiber has reached the same barrier.
 * Blocks run one after the other.  Warp shuffles are emulated through a per-block exchange
 * buffer.  Device-only instructions (f32x2 PTX, cp.async, mbarrier) have host equivalents in
 * the kernel headers themselves under `#ifndef __CUDA_ARCH__`.
 */
#ifndef CUDA_EMU_H
#define CUDA_EMU_H

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define B200_EMULATED 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace emu {
struct Fiber {
    ucontext_t ctx;
    char *stack;
    bool done;
    uint3_emu tid;
};
struct State {
    std::vector<Fiber> fibers;
    ucontext_t main_ctx;
    int current;
    dim3 grid, block;
    uint3_emu bid;
    unsigned char *dyn_smem;
    size_t dyn_smem_bytes;
    std::function<void()> body;
    uint32_t shfl_buf[1024];
    int bar_count, bar_gen;           /* block barrier */
    int wbar_count[32], wbar_gen[32]; /* per-warp barriers */
    int nbar_count[16], nbar_gen[16]; /* named barriers (bar.sync id, count) */
};
inline State &S() { static State s; return s; }
inline void trampoline()
{
    State &s = S();
    s.body();
    s.fibers[s.current].done = true;
    swapcontext(&s.fibers[s.current].ctx, &s.main_ctx);
}
inline void yield_now()
{
    State &s = S();
    swapcontext(&s.fibers[s.current].ctx, &s.main_ctx);
}
/* counting barriers: a fiber waits until every participant has arrived in this generation */
inline void block_barrier()
{
    State &s = S();
    int n = (int)(s.block.x * s.block.y * s.block.z);
    int g = s.bar_gen;
    if (++s.bar_count == n) { s.bar_count = 0; s.bar_gen++; }
    else while (s.bar_gen == g) yield_now();
}
inline void warp_barrier()
{
    State &s = S();
    int n = (int)(s.block.x * s.block.y * s.block.z);
    int w = s.current >> 5;
    int members = (w + 1) * 32 <= n ? 32 : n - w * 32;
    int g = s.wbar_gen[w];
    if (++s.wbar_count[w] == members) { s.wbar_count[w] = 0; s.wbar_gen[w]++; }
    else while (s.wbar_gen[w] == g) yield_now();
}
/* bar.sync id, count: `count` fibers meet at barrier `id` */
inline void named_barrier(int id, int count)
{
    State &s = S();
    int g = s.nbar_gen[id];
    if (++s.nbar_count[id] == count) { s.nbar_count[id] = 0; s.nbar_gen[id]++; }
    else while (s.nbar_gen[id] == g) yield_now();
}
/* bar.arrive id, count: count this fiber in and go on */
inline void named_arrive(int id, int count)
{
    State &s = S();
    if (++s.nbar_count[id] == count) { s.nbar_count[id] = 0; s.nbar_gen[id]++; }
}
/* run one block */
inline void run_block(const std::function<void()> &body)
{
    State &s = S();
    unsigned n = s.block.x * s.block.y * s.block.z;
    s.body = body;
    s.fibers.resize(n);
    s.bar_count = 0; s.bar_gen = 0;
    for (int w = 0; w < 32; ++w) { s.wbar_count[w] = 0; s.wbar_gen[w] = 0; }
    for (int b = 0; b < 16; ++b) { s.nbar_count[b] = 0; s.nbar_gen[b] = 0; }
    const size_t STK = 256 * 1024;
    for (unsigned i = 0; i < n; ++i) {
        Fiber &f = s.fibers[i];
        f.stack = (char *)malloc(STK);
        f.done = false;
        f.tid.x = i % s.block.x;
        f.tid.y = (i / s.block.x) % s.block.y;
        f.tid.z = i / (s.block.x * s.block.y);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STK;
        f.ctx.uc_link = &s.main_ctx;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    bool any = true;
    while (any) {
        any = false;
        for (unsigned i = 0; i < n; ++i) {
            if (s.fibers[i].done) continue;
            any = true;
            s.current = (int)i;
            swapcontext(&s.main_ctx, &s.fibers[i].ctx);
        }
    }
    for (unsigned i = 0; i < n; ++i) free(s.fibers[i].stack);
}
template <typename F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F &&body)
{
    State &s = S();
    s.grid = grid;
    s.block = block;
    s.dyn_smem_bytes = smem_bytes;
    s.dyn_smem = (unsigned char *)aligned_alloc(1024, ((smem_bytes + 1023) / 1024 + 1) * 1024);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                s.bid = uint3_emu{bx, by, bz};
                memset(s.dyn_smem, 0xCD, smem_bytes); /* poison: shared memory starts undefined */
                run_block(body);
            }
    free(s.dyn_smem);
}
} // namespace emu

#define threadIdx (emu::S().fibers[emu::S().current].tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().block)
#define gridDim (emu::S().grid)
#define EMU_DYN_SMEM (emu::S().dyn_smem)

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { (void)m; emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

/* every lane of every warp must call shuffles convergently (true for our kernels) */
static inline uint32_t emu_shfl_raw(uint32_t v, int src_lane_abs)
{
    emu::State &s = emu::S();
    int me = s.current;
    s.shfl_buf[me] = v;
    emu::warp_barrier();
    uint32_t r = s.shfl_buf[src_lane_abs];
    emu::warp_barrier();
    return r;
}
static inline float __shfl_sync(unsigned m, float v, int lane)
{
    (void)m;
    int base = emu::S().current & ~31;
    uint32_t u; memcpy(&u, &v, 4);
    u = emu_shfl_raw(u, base + (lane & 31));
    float r; memcpy(&r, &u, 4);
    return r;
}
static inline float __shfl_up_sync(unsigned m, float v, unsigned d)
{
    (void)m;
    int me = emu::S().current, lane = me & 31;
    int src = lane >= (int)d ? me - (int)d : me;
    uint32_t u; memcpy(&u, &v, 4);
    u = emu_shfl_raw(u, src);
    float r; memcpy(&r, &u, 4);
    return r;
}
static inline float __shfl_xor_sync(unsigned m, float v, int x)
{
    (void)m;
    int me = emu::S().current;
    uint32_t u; memcpy(&u, &v, 4);
    u = emu_shfl_raw(u, (me & ~31) | ((me ^ x) & 31));
    float r; memcpy(&r, &u, 4);
    return r;
}

/* integer shuffles, votes and reductions (misc_kernels.cuh k_counter_check) */
static inline uint32_t __shfl_sync(unsigned m, uint32_t v, int lane)
{
    (void)m;
    return emu_shfl_raw(v, (emu::S().current & ~31) + (lane & 31));
}
static inline uint32_t __shfl_down_sync(unsigned m, uint32_t v, unsigned d)
{
    (void)m;
    int me = emu::S().current, lane = me & 31;
    return emu_shfl_raw(v, lane + (int)d < 32 ? me + (int)d : me);
}
static inline uint64_t __shfl_xor_sync(unsigned m, uint64_t v, int x)
{
    (void)m;
    int src = (emu::S().current & ~31) | ((emu::S().current ^ x) & 31);
    uint32_t lo = emu_shfl_raw((uint32_t)v, src), hi = emu_shfl_raw((uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
static inline uint32_t __reduce_add_sync(unsigned m, uint32_t v)
{
    (void)m;
    emu::State &s = emu::S();
    s.shfl_buf[s.current] = v;
    emu::warp_barrier();
    uint32_t r = 0;
    for (int l = 0; l < 32; ++l) r += s.shfl_buf[(s.current & ~31) + l];
    emu::warp_barrier();
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __reduce_add_sync(m, pred ? 1u : 0u) != 0u; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift)
{
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (shift & 31));
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline uint4 __ldcs(const uint4 *p) { return *p; }
static inline unsigned __ldcs(const unsigned *p) { return *p; }
static inline void __stcs(float4 *p, float4 v) { *p = v; }

static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t s = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)((v >> (8 * (s & 7))) & 0xFF);
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __ldg(const float *p) { return *p; }
static inline float __ldcg(const float *p) { return *p; }
static inline float4 __ldcg(const float4 *p) { return *p; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline int __float2int_rn(float f) { return (int)nearbyintf(f); }
static inline unsigned __ldg(const unsigned *p) { return *p; }
static inline unsigned short __ldg(const unsigned short *p) { return *p; }
static inline uint4 __ldg(const uint4 *p) { return *p; }
static inline float2 __ldg(const float2 *p) { return *p; }
static inline float4 __ldg(const float4 *p) { return *p; }
static inline float atomicAdd(float *p, float v) { float o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicInc(unsigned *p, unsigned wrap) { unsigned o = *p; *p = (o >= wrap) ? 0u : o + 1u; return o; }

#endif
