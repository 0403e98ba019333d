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
