/*
 * b200sdr_synth.h -- bit-reproducible synthetic RTL-SDR captures (u8 interleaved I/Q).
 *
 * The reference firmware never sees RF in its committed configuration: it turns on the
 * RTL2832 test mode (reference RTL/Src/usbh_rtlsdr.c:901, :662-664) and copies whatever
 * bytes arrive on EP 0x81 into its buffer (usbh_rtlsdr.c:1068-1077).  The measurement
 * plan therefore uses synthetic captures.  This header is the one
 * definition of those captures; it is compiled by gcc (oracle, CPU reference arm) and by
 * nvcc (device generator in csrc/synth.cu) and produces IDENTICAL bytes on both, because
 * it only uses integer arithmetic, a host-computed sine table and individually rounded
 * IEEE-754 float/double operations (no fused multiply-add, no libm inside the formula).
 *
 * Kinds:
 *   COUNTER   u[i] = i mod 256                (RTL2832 test-mode stream, ingest fixture)
 *   MULTITONE 4 complex tones + noise        (spectrum chain)
 *   WBFM      FM, 75 kHz deviation, 1 kHz + 5 kHz message, +50 kHz carrier offset
 *   AM        (1 + 0.5 m) * 0.5 envelope on a +10 kHz carrier offset
 *
 * Sample rate is fixed at 2.4 MS/s (BASELINE.json configs).
 */
#ifndef B200SDR_SYNTH_H
#define B200SDR_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define B200SDR_HD __host__ __device__ __forceinline__
#else
#define B200SDR_HD static inline
#endif

/* individually rounded arithmetic: the device compiler must not contract a*b+c */
#if defined(__CUDA_ARCH__)
#define B200SDR_FMUL(a, b) __fmul_rn((a), (b))
#define B200SDR_FADD(a, b) __fadd_rn((a), (b))
#define B200SDR_DMUL(a, b) __dmul_rn((a), (b))
#define B200SDR_DADD(a, b) __dadd_rn((a), (b))
#else /* host: build with -ffp-contract=off */
#define B200SDR_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define B200SDR_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define B200SDR_DMUL(a, b) ((double)((double)(a) * (double)(b)))
#define B200SDR_DADD(a, b) ((double)((double)(a) + (double)(b)))
#endif

#define B200SDR_SYNTH_COUNTER   0u
#define B200SDR_SYNTH_MULTITONE 1u
#define B200SDR_SYNTH_WBFM      2u
#define B200SDR_SYNTH_AM        3u

#define B200SDR_SYNTH_SEED_BASE 0x5D12B200ull /* seed = base + capture index */
#define B200SDR_SYNTH_LUT_BITS  12
#define B200SDR_SYNTH_LUT_SIZE  (1u << B200SDR_SYNTH_LUT_BITS)
/* table holds LUT_SIZE + 1 entries: lut[i] = (float) sin(2 pi i / LUT_SIZE), computed on the
 * host in double precision; lut[LUT_SIZE] = lut[0] closes the interpolation interval. */

/* phase increments, cycles/sample * 2^32, fs = 2.4 MS/s (all exactly representable) */
#define B200SDR_PH_PER_HZ (4294967296.0 / 2400000.0)

B200SDR_HD uint64_t b200sdr_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* sin(2 pi phase / 2^32) by linear interpolation in the table */
B200SDR_HD float b200sdr_lut_sin(const float *lut, uint32_t phase)
{
    uint32_t idx = phase >> (32 - B200SDR_SYNTH_LUT_BITS);
    uint32_t fr = phase & ((1u << (32 - B200SDR_SYNTH_LUT_BITS)) - 1u);
    float f = B200SDR_FMUL((float)fr, 1.0f / (float)(1u << (32 - B200SDR_SYNTH_LUT_BITS)));
    float a = lut[idx], b = lut[idx + 1];
    return B200SDR_FADD(a, B200SDR_FMUL(f, B200SDR_FADD(b, -a)));
}
B200SDR_HD float b200sdr_lut_cos(const float *lut, uint32_t phase)
{
    return b200sdr_lut_sin(lut, phase + 0x40000000u);
}

/* approximately N(0,1): sum of four 16-bit uniforms (Irwin-Hall), integer until the last step */
B200SDR_HD float b200sdr_gauss(uint64_t h)
{
    int32_t s = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) +
                (int32_t)((h >> 48) & 0xFFFF) - 2 * 65535;
    /* var of one uniform on 0..65535 = (65536^2 - 1)/12 ; four of them: sigma = 37837.23 */
    return B200SDR_FMUL((float)s, 1.0f / 37837.23f);
}

B200SDR_HD uint8_t b200sdr_quantise(float s)
{
    /* u = clamp(round(127.5 + 127.5 * 0.9 * s), 0, 255) */
    float v = B200SDR_FADD(B200SDR_FADD(127.5f, B200SDR_FMUL(114.75f, s)), 0.5f);
    int32_t q = (int32_t)v; /* v >= 0 in range; truncation == floor */
    if (v < 0.0f) q = 0;
    if (q > 255) q = 255;
    return (uint8_t)q;
}

B200SDR_HD uint32_t b200sdr_phase_at(uint32_t phi0, uint32_t inc, uint64_t n)
{
    return phi0 + (uint32_t)((uint64_t)inc * n);
}

/* One complex sample n (0-based) of capture `seed` -> two bytes (I, Q). */
B200SDR_HD void b200sdr_synth_sample(const float *lut, uint32_t kind, uint64_t seed, uint64_t n, uint8_t *out_i,
                                     uint8_t *out_q)
{
    if (kind == B200SDR_SYNTH_COUNTER) {
        *out_i = (uint8_t)((2 * n) & 0xFF);
        *out_q = (uint8_t)((2 * n + 1) & 0xFF);
        return;
    }
    const float sigma = 0.01f;
    uint64_t hs = b200sdr_mix64(seed);
    float re = 0.0f, im = 0.0f;
    if (kind == B200SDR_SYNTH_MULTITONE) {
        /* f = {-600k, -150k, +37.5k, +450k} Hz ; A = {0.5, 0.25, 0.1, 0.05} */
        const uint32_t inc[4] = {0xC0000000u, 0xF0000000u, 0x04000000u, 0x30000000u};
        const float amp[4] = {0.5f, 0.25f, 0.1f, 0.05f};
        for (int k = 0; k < 4; ++k) {
            uint32_t phi0 = (uint32_t)b200sdr_mix64(hs + (uint64_t)k);
            uint32_t ph = b200sdr_phase_at(phi0, inc[k], n);
            re = B200SDR_FADD(re, B200SDR_FMUL(amp[k], b200sdr_lut_cos(lut, ph)));
            im = B200SDR_FADD(im, B200SDR_FMUL(amp[k], b200sdr_lut_sin(lut, ph)));
        }
    } else {
        /* message m = 0.5 sin(2 pi 1k t) + 0.3 sin(2 pi 5k t) */
        const uint32_t inc1 = 1789570u;  /* round(1000 * 2^32 / 2.4e6)  = 1789569.7 */
        const uint32_t inc2 = 8947849u;  /* round(5000 * 2^32 / 2.4e6)  = 8947848.5 */
        uint32_t p1 = b200sdr_phase_at((uint32_t)b200sdr_mix64(hs + 11u), inc1, n);
        uint32_t p2 = b200sdr_phase_at((uint32_t)b200sdr_mix64(hs + 12u), inc2, n);
        uint32_t phi0 = (uint32_t)b200sdr_mix64(hs + 13u);
        if (kind == B200SDR_SYNTH_WBFM) {
            /* phase deviation in cycles: fd*0.5/(2 pi f1) (1-cos p1) + fd*0.3/(2 pi f2) (1-cos p2)
             * fd = 75 kHz: 5.968310365946075 and 0.716197243913529 cycles */
            const uint32_t incc = 89478485u; /* +50 kHz: round(50000 * 2^32 / 2.4e6) = 89478485.3 */
            double d1 = B200SDR_DMUL(5.968310365946075 * 4294967296.0,
                                     (double)B200SDR_FADD(1.0f, -b200sdr_lut_cos(lut, p1)));
            double d2 = B200SDR_DMUL(0.716197243913529 * 4294967296.0,
                                     (double)B200SDR_FADD(1.0f, -b200sdr_lut_cos(lut, p2)));
            uint64_t dev = (uint64_t)B200SDR_DADD(B200SDR_DADD(d1, d2), 0.5);
            uint32_t ph = b200sdr_phase_at(phi0, incc, n) + (uint32_t)dev;
            re = B200SDR_FMUL(0.8f, b200sdr_lut_cos(lut, ph));
            im = B200SDR_FMUL(0.8f, b200sdr_lut_sin(lut, ph));
        } else { /* AM */
            const uint32_t incc = 17895697u; /* +10 kHz: round(10000 * 2^32 / 2.4e6) = 17895697.07 */
            float m = B200SDR_FADD(B200SDR_FMUL(0.5f, b200sdr_lut_sin(lut, p1)),
                                   B200SDR_FMUL(0.3f, b200sdr_lut_sin(lut, p2)));
            float env = B200SDR_FMUL(0.5f, B200SDR_FADD(1.0f, B200SDR_FMUL(0.5f, m)));
            uint32_t ph = b200sdr_phase_at(phi0, incc, n);
            re = B200SDR_FMUL(env, b200sdr_lut_cos(lut, ph));
            im = B200SDR_FMUL(env, b200sdr_lut_sin(lut, ph));
        }
    }
    uint64_t hn = b200sdr_mix64(hs ^ (n * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
    re = B200SDR_FADD(re, B200SDR_FMUL(sigma, b200sdr_gauss(hn)));
    im = B200SDR_FADD(im, B200SDR_FMUL(sigma, b200sdr_gauss(b200sdr_mix64(hn))));
    *out_i = b200sdr_quantise(re);
    *out_q = b200sdr_quantise(im);
}

#endif /* B200SDR_SYNTH_H */
