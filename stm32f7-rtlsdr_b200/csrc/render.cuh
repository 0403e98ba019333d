/*
 * render.cuh -- presentation step after the averaging: the averaged
 * power spectrum as a 480 x 272 ARGB8888 bar plot, the format of the LCD layer the firmware's
 * sample buffer aliases (layer 1 at LCD_FB_START_ADDRESS + 480*272*4, src/main.c:100-109;
 * RK043FN48H 480 x 272; colours 0xAARRGGBB as in stm32746g_discovery_lcd.h:123-134).
 * The firmware only draws a test circle there (main.c:111-114); the README's goal is a spectrum
 * analyser view (README.md:7-10).  Definition followed: oracle/golden.c gold_render_spectrum().
 *
 * Everything after the power values is integer arithmetic and float comparisons against a
 * host-built threshold table, so the image is bit-exact against the oracle:
 *   column c  <- max power over fft-shifted bins [floor(c 1024/480), floor((c+1) 1024/480))
 *   height    <- number of thresholds T[h] <= power, T[h] = 10^((db_min + (db_max-db_min) h/271)/10)
 *   pixel     <- colour ramp blue-cyan-green-yellow-red of its own height if below the bar top
 */
#ifndef B200_RENDER_CUH
#define B200_RENDER_CUH

#include <stdint.h>

#define B200_LCD_W 480
#define B200_LCD_H 272

#if defined(__CUDACC__) && !defined(B200_EMULATED)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD static inline
#endif

B200_HD uint32_t b200_ramp_argb(int y) /* y = height above the bottom row, 0..271 */
{
    const int i = (y * 255) / (B200_LCD_H - 1);
    const int seg = i >> 6, t = (i & 63) << 2;
    int r, g, b;
    if (seg == 0) { r = 0; g = t; b = 255; }
    else if (seg == 1) { r = 0; g = 255; b = 255 - t; }
    else if (seg == 2) { r = t; g = 255; b = 0; }
    else { r = 255; g = 255 - t; b = 0; }
    return 0xFF000000u | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
}

/* host side: dB thresholds of the 272 LCD rows, T[h] = 10^((db_min + (db_max - db_min) h / 271) / 10), computed in
 * double and rounded once (shared by the C ABI and the test-suite's host emulation of the kernels) */
#include <math.h>
static inline void b200_fill_thresholds(float *thr272, float db_min, float db_max) /* host function */
{
    for (int h = 0; h < B200_LCD_H; ++h)
        thr272[h] = (float)pow(10.0, ((double)db_min + ((double)db_max - (double)db_min) * (double)h / (B200_LCD_H - 1.0)) / 10.0);
}

/* one CTA of 480 threads per spectrum; `scale` multiplies the stored values first (1/frames for
 * the streaming accumulator, which holds the sum) */
__global__ void __launch_bounds__(B200_LCD_W) k_render_spectrum(const float *__restrict__ spectra, float scale,
                                                              const float *__restrict__ thresholds,
                                                              uint32_t *__restrict__ argb)
{
    __shared__ float s_thr[B200_LCD_H];
    const int c = (int)threadIdx.x;
    const float *spec = spectra + (uint64_t)blockIdx.x * 1024u;
    uint32_t *img = argb + (uint64_t)blockIdx.x * (B200_LCD_W * B200_LCD_H);
    if (c < B200_LCD_H) s_thr[c] = thresholds[c];
    __syncthreads();
    const int s0 = (c * 1024) / B200_LCD_W, s1 = ((c + 1) * 1024) / B200_LCD_W;
    float v = 0.0f;
    for (int s = s0; s < s1; ++s) {
        const float pwr = spec[(s + 512) & 1023] * scale; /* one rounded multiply, as the host getter does */
        v = pwr > v ? pwr : v;
    }
    /* thresholds ascend: binary search for the count of T[h] <= v */
    int lo = 0, hi = B200_LCD_H;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_thr[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    const int height = lo;
    for (int r = 0; r < B200_LCD_H; ++r) {
        const int y = B200_LCD_H - 1 - r;
        img[r * B200_LCD_W + c] = y < height ? b200_ramp_argb(y) : 0xFF000000u;
    }
}

/* Waterfall view (oracle/golden.c gold_render_waterfall): one CTA per image row, one thread per pixel;
 * row r shows spectrum r in the colour of the bar height its power would reach, rows >= n_rows are black. */
__global__ void __launch_bounds__(B200_LCD_W) k_render_waterfall(const float *__restrict__ spectra, uint32_t n_rows,
                                                               float scale, const float *__restrict__ thresholds,
                                                               uint32_t *__restrict__ argb)
{
    __shared__ float s_thr[B200_LCD_H];
    const int c = (int)threadIdx.x, r = (int)blockIdx.x;
    if (c < B200_LCD_H) s_thr[c] = thresholds[c];
    __syncthreads();
    uint32_t px = 0xFF000000u;
    if ((uint32_t)r < n_rows) {
        const float *spec = spectra + (uint64_t)r * 1024u;
        const int s0 = (c * 1024) / B200_LCD_W, s1 = ((c + 1) * 1024) / B200_LCD_W;
        float v = 0.0f;
        for (int s = s0; s < s1; ++s) {
            const float pwr = spec[(s + 512) & 1023] * scale;
            v = pwr > v ? pwr : v;
        }
        int lo = 0, hi = B200_LCD_H;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_thr[mid] <= v) lo = mid + 1;
            else hi = mid;
        }
        if (lo > 0) px = b200_ramp_argb(lo - 1);
    }
    argb[r * B200_LCD_W + c] = px;
}

#endif
