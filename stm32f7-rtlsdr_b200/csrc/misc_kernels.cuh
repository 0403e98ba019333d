/*
 * misc_kernels.cuh -- K2 (stand-alone u8 -> cf32 (+window) conversion) and the on-device
 * synthetic capture generator.
 */
#ifndef B200_MISC_KERNELS_CUH
#define B200_MISC_KERNELS_CUH

#include "../../include/b200sdr_synth.h"
#include "cplx2.cuh"

/* K2: out[2n] = (I_n - 127.5) w[n mod 1024], out[2n+1] = (Q_n - 127.5) w[n mod 1024].
 * A thread converts four 32-bit words (2 complex samples each) taken 256 words apart, so every warp
 * load is 128 contiguous bytes and every warp store 512 contiguous bytes (whole sectors, no partial
 * writes); streaming cache hints: the data is touched once.  With window == nullptr the conversion
 * is exact and unscaled (bit-exact parity row of SURVEY.md section 8a).  `n_words` = len / 4. */
__global__ void __launch_bounds__(256) k_convert_cf32(const uint32_t *__restrict__ in, float4 *__restrict__ out,
                                                      uint64_t n_words, const float *__restrict__ window)
{
    const uint64_t base = (uint64_t)blockIdx.x * 1024u + threadIdx.x;
    uint32_t wd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t j = base + 256u * k;
        wd[k] = j < n_words ? __ldcs(in + j) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t j = base + 256u * k;
        if (j >= n_words) continue;
        float w0 = 1.0f, w1 = 1.0f;
        if (window) {
            const float2 ww = __ldg(reinterpret_cast<const float2 *>(window + ((j * 2u) & 1023u)));
            w0 = ww.x;
            w1 = ww.y;
        }
        float4 o;
        /* __fmul_rn: keep the product a plain rounded multiply (nothing to contract with) */
        o.x = __fmul_rn(b200_u8_to_f32(wd[k], 0), w0);
        o.y = __fmul_rn(b200_u8_to_f32(wd[k], 1), w0);
        o.z = __fmul_rn(b200_u8_to_f32(wd[k], 2), w1);
        o.w = __fmul_rn(b200_u8_to_f32(wd[k], 3), w1);
        __stcs(out + j, o);
    }
}

/* K0: test-mode counter check.  The firmware switches the RTL2832 into test mode before it starts the bulk
 * transfers (RTLSDR_set_test_mode(phost, 1), RTL/Src/usbh_rtlsdr.c:901 and :660-662), so the byte stream that
 * reaches the buffer boundary is the dongle's 8-bit counter: u[i] = u[i-1] + 1 (mod 256).  A byte that is not
 * its predecessor + 1 marks lost samples (what rtl_test counts).  This kernel counts those breaks and finds the
 * first one, per capture; the only HBM-bound chain of the path next to K2: 1 byte read per byte, nothing written.
 *   a break at index i (1 <= i < len)  <=>  u[i] != (u[i-1] + 1) & 0xff;  i = 0 breaks iff expect_first >= 0 and
 *   u[0] != expect_first (the value that continues the previous block).
 * A warp takes 128 consecutive 16-byte vectors as four loads of 512 contiguous bytes; inside a
 * word the four byte lanes are compared at once: s = the word shifted down one byte with the next word's first
 * byte on top (funnel shift) must equal the word with every byte incremented.  The next vector's first word comes
 * from the neighbouring lane (lane 31: from lane 0 of the warp's next load).  `n_vec` = capture length / 16, `n_words` = length / 4 (the tail of a
 * length that is not a multiple of 16 is handled word by word by the last block). */
struct CounterStreamState {      /* carried from block to block by the streaming path (one capture) */
    unsigned long long n_breaks, first_break; /* first_break: absolute byte index in the stream, ~0 = none */
    uint32_t expect;             /* value the next block's first byte must have; > 255: nothing known yet */
    uint32_t pad;
};
struct CounterParams {
    const uint32_t *in;      /* capture c at in + c * stride_words; 4-byte aligned (batches: 16-byte aligned) */
    uint64_t stride_words, n_words;
    int32_t expect_first;    /* -1: the first byte is not checked (ignored when `stream` is set) */
    unsigned long long *n_breaks, *first_break; /* [capture]; preset to 0 and ~0 */
    CounterStreamState *stream; /* optional, one capture: results are accumulated here with `pos_base` added,
                                   the first byte is checked against stream->expect, and expect is updated */
    uint64_t pos_base;       /* absolute byte index of in[0] in the stream */
};

/* bytes of the result are non-zero where the successor of a byte of `w` is not that byte + 1.  The per-byte
 * increment is done inside the 32-bit word without carries between the lanes (low seven bits added, top bit
 * xor-ed back): five integer instructions per word -- the byte-SIMD intrinsics (__vsub4, __vcmpne4) are emulated
 * with three times as many on this architecture and made the kernel issue-bound at 55 % of the HBM roofline. */
__device__ __forceinline__ uint32_t b200_counter_diff(uint32_t w, uint32_t next)
{
    const uint32_t s = __funnelshift_r(w, next, 8);                                   /* bytes (b1, b2, b3, next.b0) */
    const uint32_t w1 = ((w & 0x7f7f7f7fu) + 0x01010101u) ^ (w & 0x80808080u);         /* (b0+1, b1+1, b2+1, b3+1) mod 256 */
    return s ^ w1;
}
/* the rare path: count the non-zero bytes of `ne` (word `word_idx` of the capture) and note the first */
__device__ __forceinline__ void b200_counter_note(uint32_t ne, uint64_t word_idx, uint32_t &count, uint64_t &first)
{
    if (ne == 0u) return;
    const uint32_t flags = (((ne & 0x7f7f7f7fu) + 0x7f7f7f7fu) | ne) & 0x80808080u;   /* 0x80 per non-zero byte */
    count += (uint32_t)__popc(flags);
    const uint64_t pos = word_idx * 4u + (uint32_t)((__ffs((int)flags) - 1) >> 3) + 1u; /* index of the offending byte */
    if (pos < first) first = pos;
}

__global__ void __launch_bounds__(256) k_counter_check(CounterParams p)
{
    const uint32_t c = blockIdx.y;
    const uint32_t *in0 = p.in + (uint64_t)c * p.stride_words;
    /* words before the first 16-byte boundary (streaming blocks start on any word), whole vectors, tail words */
    uint64_t head = ((16u - (uint32_t)((uintptr_t)in0 & 15u)) & 15u) >> 2;
    if (head > p.n_words) head = p.n_words;
    const uint32_t *in = in0 + head;
    const uint64_t rest = p.n_words - head, n_vec = rest / 4u;
    const int lane = (int)(threadIdx.x & 31u);
    uint32_t count = 0;
    uint64_t first = ~0ull;
    /* a warp owns 128 consecutive vectors (2 KB) as four 512-byte loads, so the successor of lane 31's vector is
     * lane 0's vector of the warp's next load, and only the vector after the warp's last one is an extra
     * (broadcast) read.  32-bit vector indices: the host keeps captures below 2^32 vectors. */
    const uint32_t nv = (uint32_t)n_vec;
    const uint32_t wbase = blockIdx.x * 1024u + (threadIdx.x >> 5) * 128u;
    const uint4 *vec = reinterpret_cast<const uint4 *>(in);
    if (wbase + 128u < nv) { /* warp-uniform: all 128 vectors and the one after them exist */
        uint4 vv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) vv[k] = __ldcs(vec + (wbase + 32u * k + lane)); /* all loads in flight first */
        const uint32_t after = __ldg(in + (uint64_t)(wbase + 128u) * 4u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = vv[k];
            const uint32_t up = __shfl_down_sync(0xffffffffu, v.x, 1);
            const uint32_t wrap = k < 3 ? __shfl_sync(0xffffffffu, vv[k < 3 ? k + 1 : 3].x, 0) : after;
            const uint32_t next = lane == 31 ? wrap : up;
            const uint32_t e0 = b200_counter_diff(v.x, v.y), e1 = b200_counter_diff(v.y, v.z);
            const uint32_t e2 = b200_counter_diff(v.z, v.w), e3 = b200_counter_diff(v.w, next);
            if (e0 | e1 | e2 | e3) {
                const uint64_t w0 = head + 4u * (uint64_t)(wbase + 32u * k + lane);
                b200_counter_note(e0, w0, count, first);
                b200_counter_note(e1, w0 + 1u, count, first);
                b200_counter_note(e2, w0 + 2u, count, first);
                b200_counter_note(e3, w0 + 3u, count, first);
            }
        }
    } else if (wbase < nv) { /* the last warp(s) of a capture: bounds per vector */
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const uint64_t j = (uint64_t)wbase + 32u * k + lane;
            const uint4 v = j < n_vec ? __ldcs(vec + j) : make_uint4(0u, 0u, 0u, 0u);
            /* the neighbouring lane has the next vector's first word; lane 31 and the last vector read it;
             * the very last byte of a capture has no successor, so it is given the one it expects */
            uint32_t next = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 31 || j + 1 >= n_vec) next = (j + 1) * 4u < rest ? __ldg(in + (j + 1) * 4u) : (v.w >> 24) + 1u;
            if (j < n_vec) {
                b200_counter_note(b200_counter_diff(v.x, v.y), head + 4u * j, count, first);
                b200_counter_note(b200_counter_diff(v.y, v.z), head + 4u * j + 1u, count, first);
                b200_counter_note(b200_counter_diff(v.z, v.w), head + 4u * j + 2u, count, first);
                b200_counter_note(b200_counter_diff(v.w, next), head + 4u * j + 3u, count, first);
            }
        }
    }
    /* head words, words past the last whole vector, the first byte and the carried expectation: one thread */
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.n_words > 0) {
        for (int part = 0; part < 2; ++part) {
            const uint64_t lo = part ? head + n_vec * 4u : 0u, hi = part ? p.n_words : head;
            for (uint64_t w = lo; w < hi; ++w)
                b200_counter_note(b200_counter_diff(in0[w], w + 1 < p.n_words ? in0[w + 1] : (in0[w] >> 24) + 1u), w, count, first);
        }
        int32_t expect = p.expect_first;
        if (p.stream) expect = p.stream->expect > 255u ? -1 : (int32_t)p.stream->expect;
        if (expect >= 0 && (in0[0] & 0xffu) != (uint32_t)expect) {
            count += 1;
            first = 0;
        }
        if (p.stream) p.stream->expect = ((in0[p.n_words - 1] >> 24) + 1u) & 0xffu;
    }
    if (__any_sync(0xffffffffu, first != ~0ull)) { /* rare: a clean stream never gets here */
        count = __reduce_add_sync(0xffffffffu, count);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, first, o);
            if (other < first) first = other;
        }
        if (lane == 0) {
            if (p.stream) {
                atomicAdd(&p.stream->n_breaks, (unsigned long long)count);
                atomicMin(&p.stream->first_break, (unsigned long long)(p.pos_base + first));
            } else {
                atomicAdd(p.n_breaks + c, (unsigned long long)count);
                atomicMin(p.first_break + c, (unsigned long long)first);
            }
        }
    }
}

/* synthetic captures: one thread per 8 complex samples (16 bytes) */
__global__ void __launch_bounds__(256) k_synth(uint4 *__restrict__ out, uint64_t groups_per_capture,
                                               uint64_t capture_stride16, uint32_t kind, uint64_t first_capture,
                                               const float *__restrict__ lut)
{
    const uint64_t g = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (g >= groups_per_capture) return;
    const uint32_t c = blockIdx.y;
    const uint64_t seed = B200SDR_SYNTH_SEED_BASE + first_capture + c;
    uint32_t wd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint8_t i0, q0, i1, q1;
        b200sdr_synth_sample(lut, kind, seed, g * 8u + 2u * k, &i0, &q0);
        b200sdr_synth_sample(lut, kind, seed, g * 8u + 2u * k + 1u, &i1, &q1);
        wd[k] = (uint32_t)i0 | ((uint32_t)q0 << 8) | ((uint32_t)i1 << 16) | ((uint32_t)q1 << 24);
    }
    out[(uint64_t)c * capture_stride16 + g] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
}

#endif
