/*
 * spectrum.cuh -- kernel K3: u8 I/Q -> window -> 1024-point FFT -> |X|^2 -> averaged spectrum,
 * fused, one pass over the input bytes.
 *
 * Reference anchor: this is the consumer of `CommItf.buff` the firmware never wrote
 * (README.md:31-32 "next task"; the commented poll at src/main.c:76-79); the intended MCU shape
 * is arm_cfft_f32(1024) + arm_cmplx_mag_squared_f32 (CMSIS/core/arm_math.h:2149, :4693).
 * Definition followed: oracle/golden.c gold_spectrum().
 *
 * Decomposition: 1024 = 32 x 32 "four-step" FFT, ONE WARP PER FRAME, 32 points per lane:
 *   n = t + 32 j   (t = lane, j = 0..31)       k = k1 + 32 k2
 *   pass 1 (lane t):  Y[k1] = sum_j w[n] x[n] W32^{j k1}      in registers (fft32.cuh)
 *   twiddle:          Y[k1] *= W1024^{t k1}                    folded into pass 2's first stage
 *   transpose:        lane t -> lane k1 through a warp-private 32x34 shared tile
 *   pass 2 (lane k1): X[k1+32 k2] = sum_t Y_t[k1] W32^{t k2}   in registers
 *   power:            acc[k2] += |X|^2                          32 bins per lane, in registers
 * The 32 window values and 32 twiddles a lane needs never change, so they live in registers for the
 * whole kernel (96 of the 250 registers; 2 CTAs of 4 warps per SM), loaded straight from a 12 KB
 * per-lane table the host lays out once (plan.h fill_lane_consts): the only shared-memory traffic per
 * frame is the transpose.  DESIGN.md 5.1 has the A/B measurements behind these choices.
 *
 * A warp walks `frames_per_warp` consecutive frames (hop 512) and keeps the 32 x 32 bin sums in
 * registers; the warps of a CTA are then added in a fixed order and one 1024-float partial per
 * CTA goes to a workspace that k_spectrum_finalize (or the split-capture exchange kernel) sums in a
 * fixed order -> deterministic results, no float atomics.
 *
 * Input access: lane t reads the 2-byte sample n = t + 32 j straight from global memory
 * (a warp reads 64 contiguous bytes per j; both halves of every 128-byte line are used by
 * consecutive j, the second from L1), the loads of frame m+1 are issued before the FFT of frame m.
 * Every input byte is fetched from HBM once (the 50 % overlap re-read hits L2).
 *
 * Roofline: 2 bytes per new complex sample vs 482 packed + 64 scalar FP instructions per frame
 * (512 new samples) -> bound by the scheduler issue port / FP32 pipe, not by HBM.
 */
#ifndef B200_SPECTRUM_CUH
#define B200_SPECTRUM_CUH

#include "fft32.cuh"

#ifndef B200_SPEC_MINB
#define B200_SPEC_MINB 2 /* CTAs per SM the register allocation is tuned for (8 warps/SM, up to 255 registers) */
#endif
#define B200_SPEC_WARPS 4
#define B200_SPEC_THREADS (32 * B200_SPEC_WARPS)
#define B200_SPEC_XP 34 /* transpose tile row pitch in complex elements: rows are 16-byte aligned; the 64-bit
                           stores and the 128-bit loads are both conflict-free */
#define B200_SPEC_LANE_CONSTS 96 /* floats per lane: 16 twiddle pairs (float4 each) + 8 float4 of window values */

/* shared memory: the warps' transpose tiles, re-used for the CTA reduction at the end */
#define B200_SPEC_SMEM_BYTES (B200_SPEC_WARPS * 32 * B200_SPEC_XP * 8)

#ifdef B200_EMULATED
#define B200_DYN_SMEM(name) unsigned char *name = EMU_DYN_SMEM
#else
#define B200_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

struct SpectrumParams {
    const uint8_t *iq;        /* capture c starts at iq + c * capture_stride                     */
    uint64_t capture_stride;  /* bytes                                                           */
    uint32_t frames;          /* frames per capture (all captures equal length)                  */
    uint32_t frames_per_warp; /* consecutive frames walked by one warp                           */
    const float4 *lane_consts;/* [32 lanes][24 float4], plan.h fill_lane_consts                  */
    float *partials;          /* [capture][ctas_per_capture][1024]                               */
    uint32_t ctas_per_capture;
    float ema_log2_decay;     /* log2(1 - beta), EMA only                                        */
    float ema_beta;
};

template <bool EMA>
__global__ void __launch_bounds__(B200_SPEC_THREADS, B200_SPEC_MINB) k_spectrum(SpectrumParams p)
{
    B200_DYN_SMEM(smem);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    c2 *s_xp = reinterpret_cast<c2 *>(smem) + warp * (32 * B200_SPEC_XP);

    /* this lane's constants -> registers, in the order the two fused first stages consume them:
     *   r_tw[i]   = (W^{e t}, W^{(e+16) t}),  e = bitrev5(2 i): butterfly i of pass 2's first stage
     *   r_win[i2] = (w_a, w_b) of butterflies 2 i2 and 2 i2 + 1 of pass 1's first stage,
     *               w_a = w[t + 32 e], w_b = w[t + 32 (e + 16)] */
    float4 r_tw[16], r_win[8];
    {
        const float4 *lc = p.lane_consts + lane * (B200_SPEC_LANE_CONSTS / 4);
#pragma unroll
        for (int i = 0; i < 16; ++i) r_tw[i] = __ldg(lc + i);
#pragma unroll
        for (int i = 0; i < 8; ++i) r_win[i] = __ldg(lc + 16 + i);
    }

    const uint32_t capture = blockIdx.y;
    const uint32_t warp_global = blockIdx.x * B200_SPEC_WARPS + (uint32_t)warp;
    const uint32_t m_begin = warp_global * p.frames_per_warp;
    uint32_t m_end = m_begin + p.frames_per_warp;
    if (m_end > p.frames) m_end = p.frames;

    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;

    const uint8_t *cap = p.iq + (uint64_t)capture * p.capture_stride;

    /* software pipeline: the 32 two-byte loads of frame m+1 are issued before the FFT of frame m,
     * so no warp waits on HBM/L2 latency with only two warps per scheduler resident */
    const unsigned short *src0 = reinterpret_cast<const unsigned short *>(cap) + (uint32_t)lane;
    uint32_t raw[32];
    if (m_begin < m_end) {
#pragma unroll
        for (int j = 0; j < 32; ++j) raw[j] = (uint32_t)__ldg(src0 + (uint64_t)m_begin * 512u + 32 * j);
    }
    for (uint32_t m = m_begin; m < m_end; ++m) {
        c2 v[32];
        /* convert, window and first DIT stage of pass 1 in one go: butterfly i takes samples
         * e = bitrev5(2 i) and e + 16:  X = a wa + b wb,  Y = a wa - b wb  (3 packed ops) */
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) {
            const float4 w4 = r_win[i2];
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                const int i = 2 * i2 + ii, e = b200_bitrev5(2 * i);
                const c2 pa = c2_scale(c2_from_u8_lo(raw[e]), wv[2 * ii]);
                const c2 b = c2_from_u8_lo(raw[e + 16]);
                v[2 * i] = c2_fma_s(b, wv[2 * ii + 1], pa);
                v[2 * i + 1] = c2_fma_s(b, -wv[2 * ii + 1], pa);
            }
        }
        if (m + 1 < m_end) {
            const unsigned short *src = src0 + (uint64_t)(m + 1) * 512u;
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = (uint32_t)__ldg(src + 32 * j);
        }
        b200_stage_k<4, 0>::run(v);
        b200_stage_k<8, 0>::run(v);
        b200_stage_k<16, 0>::run(v);
        b200_stage_k<32, 0>::run(v); /* v[k1] = Y[k1] */
        /* transpose through the warp-private tile: row = k1 (the reader's lane), column = source lane t.
         * The barrier that keeps this frame's stores behind the previous frame's loads sits HERE, a whole
         * pass later than those loads, so it never waits on them. */
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) s_xp[k1 * B200_SPEC_XP + lane] = v[k1];
        __syncwarp();
#ifdef B200_PACKED_MATH
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
            const float4 q = *reinterpret_cast<const float4 *>(s_xp + lane * B200_SPEC_XP + t);
            v[b200_bitrev5(t)] = c2_make(q.x, q.y);
            v[b200_bitrev5(t + 1)] = c2_make(q.z, q.w);
        }
#else
#pragma unroll
        for (int t = 0; t < 32; ++t) v[b200_bitrev5(t)] = s_xp[lane * B200_SPEC_XP + t];
#endif
        /* first DIT stage of pass 2 with the four-step twiddles folded in: slots (2i, 2i+1) hold the
         * elements e = bitrev5(2i) < 16 and e + 16 of this lane's column; X = Ta a + Tb b, Y = Ta a - Tb b
         * = 2 (Ta a) - X: 5 packed ops instead of 2 + 2 (twiddles) + 2 (butterfly) */
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = b200_bitrev5(2 * i);
            const float4 tw = r_tw[i];
            const c2 pa = (e == 0) ? v[2 * i] : c2_cmul(v[2 * i], tw.x, tw.y);
            const c2 x = c2_cfma(v[2 * i + 1], tw.z, tw.w, pa);
            v[2 * i + 1] = c2_two_a_minus(pa, x);
            v[2 * i] = x;
        }
        b200_stage_k<4, 0>::run(v);
        b200_stage_k<8, 0>::run(v);
        b200_stage_k<16, 0>::run(v);
        b200_stage_k<32, 0>::run(v); /* v[k2] = X[lane + 32 k2] */
        if (EMA) {
            float wgt = p.ema_beta * exp2f((float)(p.frames - 1u - m) * p.ema_log2_decay);
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) acc[k2] = fmaf(c2_norm_acc(v[k2], 0.0f), wgt, acc[k2]);
        } else {
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) acc[k2] = c2_norm_acc(v[k2], acc[k2]);
        }
    }

    /* fixed-order reduction over the CTA's warps, then one partial per CTA */
    __syncthreads();
    float *s_red = reinterpret_cast<float *>(smem); /* [warp][1024], pitch 32*XP*2 floats */
    float *mine = s_red + warp * (32 * B200_SPEC_XP * 2);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) mine[k2 * 32 + lane] = acc[k2]; /* bin k = lane + 32 k2 */
    __syncthreads();
    float *out = p.partials + ((uint64_t)capture * p.ctas_per_capture + blockIdx.x) * 1024u;
    for (int k = tid; k < 1024; k += B200_SPEC_THREADS) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < B200_SPEC_WARPS; ++w) s += s_red[w * (32 * B200_SPEC_XP * 2) + k];
        out[k] = s;
    }
}

/* out[c][k] = scale * sum_i partials[c][i][k] + carry_scale * carry[k]  (fixed order) */
__global__ void __launch_bounds__(256) k_spectrum_finalize(const float *partials, uint32_t ctas_per_capture,
                                                           float scale, const float *carry, float carry_scale,
                                                           float *out)
{
    const uint32_t capture = blockIdx.y;
    const uint32_t k = blockIdx.x * 256u + threadIdx.x;
    const float *src = partials + (uint64_t)capture * ctas_per_capture * 1024u + k;
    float s = 0.0f;
    for (uint32_t i = 0; i < ctas_per_capture; ++i) s += src[(uint64_t)i * 1024u];
    float r = s * scale;
    if (carry) r = fmaf(carry[(uint64_t)capture * 1024u + k], carry_scale, r);
    out[(uint64_t)capture * 1024u + k] = r;
}

#endif
