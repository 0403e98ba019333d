/*
 * misc_kernels.cuh -- K2 (stand-alone u8 -> cf32 (+window) conversion) and the on-device
 * synthetic capture generator.
 */
#ifndef B200_MISC_KERNELS_CUH
#define B200_MISC_KERNELS_CUH

#include "../../include/b200sdr_synth.h"
#include "cplx2.cuh"

/* K2: out[2n] = (I_n - 127.5) w[n mod 1024], out[2n+1] = (Q_n - 127.5) w[n mod 1024].
 * One thread per 16 input bytes (8 complex samples): one 16-byte load, four 16-byte stores,
 * fully coalesced.  With window == nullptr the conversion is exact and unscaled (bit-exact parity
 * row of SURVEY.md section 8a).  `n16` = number of whole 16-byte groups; the host handles no tail
 * because the staging buffers are padded to 16 bytes. */
__global__ void __launch_bounds__(256) k_convert_cf32(const uint4 *__restrict__ in, float4 *__restrict__ out,
                                                      uint64_t n16, const float *__restrict__ window)
{
    const uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (i >= n16) return;
    const uint4 r = in[i];
    float w[8];
    if (window) {
        const float4 *wp = reinterpret_cast<const float4 *>(window + ((i * 8u) & 1023u));
        const float4 a = __ldg(wp), b = __ldg(wp + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 1.0f;
    }
    const uint32_t wd[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float4 o;
        /* __fmul_rn: keep the product a plain rounded multiply (nothing to contract with) */
        o.x = __fmul_rn(b200_u8_to_f32(wd[k], 0), w[2 * k]);
        o.y = __fmul_rn(b200_u8_to_f32(wd[k], 1), w[2 * k]);
        o.z = __fmul_rn(b200_u8_to_f32(wd[k], 2), w[2 * k + 1]);
        o.w = __fmul_rn(b200_u8_to_f32(wd[k], 3), w[2 * k + 1]);
        out[i * 4u + k] = o;
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
