/*
 * spec_tc_spike.cu -- an experimental tensor-core spectrum engine (tools/spec_tc_spike.cuh) next to the FP32 engine (csrc/spectrum.cuh) on
 * the same bytes, outside the library: (1) the raw pass-1 accumulators against a float64 DFT-32 of the same samples,
 * (2) both engines against a float64 reference of a short capture, (3) both engines timed on a resident batch.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build/spec_tc_spike tools/spec_tc_spike.cu
 *   build/spec_tc_spike [captures=64] [capture_bytes=48000000]
 */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include <cmath>
#include <cuda_runtime.h>

#include "spec_tc_spike.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static void synth(std::vector<uint8_t> &iq, uint32_t seed)
{
    const size_t n = iq.size() / 2;
    uint32_t s = seed * 2654435761u + 12345u;
    const double f[3] = {0.0712, -0.2033, 0.3391}, amp[3] = {60.0, 20.0, 2.0};
    std::vector<float> tab(1 << 16);
    for (size_t i = 0; i < tab.size(); ++i) tab[i] = (float)std::sin(2.0 * b200::kPi * i / 65536.0);
    uint32_t ph[3] = {0, 0, 0}, dph[3];
    for (int t = 0; t < 3; ++t) dph[t] = (uint32_t)(int64_t)std::llround(f[t] * 4294967296.0);
    for (size_t i = 0; i < n; ++i) {
        float re = 127.5f, im = 127.5f;
        for (int t = 0; t < 3; ++t) {
            re += (float)amp[t] * tab[((ph[t] >> 16) + 16384) & 65535];
            im += (float)amp[t] * tab[ph[t] >> 16];
            ph[t] += dph[t];
        }
        s = s * 1664525u + 1013904223u;
        re += ((s >> 8) & 0xffff) / 65536.0f * 3.0f - 1.5f;
        s = s * 1664525u + 1013904223u;
        im += ((s >> 8) & 0xffff) / 65536.0f * 3.0f - 1.5f;
        int a = (int)std::floor(re + 0.5f), b = (int)std::floor(im + 0.5f);
        iq[2 * i] = (uint8_t)(a < 0 ? 0 : a > 255 ? 255 : a);
        iq[2 * i + 1] = (uint8_t)(b < 0 ? 0 : b > 255 ? 255 : b);
    }
}

static void ref_fft(std::vector<std::complex<double>> &x)
{
    const int n = (int)x.size();
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(x[i], x[j]);
    }
    for (int len = 2; len <= n; len <<= 1)
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; ++k) {
                const double a = -2.0 * b200::kPi * k / len;
                const std::complex<double> w(std::cos(a), std::sin(a)), u = x[i + k], v = x[i + k + len / 2] * w;
                x[i + k] = u + v;
                x[i + k + len / 2] = u - v;
            }
}

struct Dev {
    float *window, *partials, *out32, *outtc, *dbg;
    float2 *twiddle;
    uint8_t *image;
    uint32_t *counter, *error;
    unsigned long long *prof;
    size_t partial_floats;
    int sm;
};

static float run_fp32(Dev &d, const uint8_t *iq, uint32_t n_captures, uint64_t len, int reps)
{
    b200::SpectrumPlan pl = b200::plan_spectrum(len, n_captures, (uint32_t)d.sm);
    SpectrumParams p{};
    p.iq = iq; p.capture_stride = len; p.frames = pl.frames; p.frames_per_warp = pl.frames_per_warp;
    p.window = d.window; p.twiddle = d.twiddle; p.partials = d.partials; p.units_per_capture = pl.units_per_capture;
    p.total_units = (uint32_t)pl.total_units; p.unit_counter = d.counter;
    if ((size_t)pl.total_units * 1024 > d.partial_floats) { printf("partials too small\n"); exit(1); }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        k_spectrum<false><<<pl.grid, B200_SPEC_THREADS, B200_SPEC_SMEM_BYTES>>>(p);
        k_spectrum_finalize<<<dim3(4, n_captures), 256>>>(d.partials, pl.units_per_capture, 1.0f / pl.frames, nullptr, 0.0f, d.out32);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static float run_tc(Dev &d, const uint8_t *iq, uint32_t n_captures, uint64_t len, int reps, bool dbg, uint32_t flags = 0)
{
    b200::SpectrumTcPlan pl = b200::plan_spectrum_tc(len, n_captures, (uint32_t)d.sm);
    SpectrumTcParams p{};
    p.iq = iq; p.capture_stride = len; p.frames = pl.frames; p.tiles_per_unit = pl.tiles_per_unit;
    p.units_per_capture = pl.units_per_capture; p.total_units = (uint32_t)pl.total_units; p.twiddle = d.twiddle;
    p.b_image = d.image; p.partials = d.partials; p.unit_counter = d.counter; p.error = d.error;
    p.dbg_y0 = dbg ? d.dbg : nullptr;
    p.dbg_flags = flags;
    p.dbg_prof = d.prof;
    CK(cudaMemset(d.prof, 0, 64));
    if ((size_t)pl.total_units * 1024 > d.partial_floats) { printf("partials too small\n"); exit(1); }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        k_spectrum_tc<false><<<pl.grid, B200_STC_THREADS, B200_STC_SMEM_BYTES>>>(p);
        k_spectrum_finalize<<<dim3(4, n_captures), 256>>>(d.partials, pl.units_per_capture, 1.0f / pl.frames, nullptr, 0.0f, d.outtc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    uint32_t err = 0;
    CK(cudaMemcpy(&err, d.error, 4, cudaMemcpyDeviceToHost));
    if (err) { printf("k_spectrum_tc: error word %u (bounded wait expired)\n", err); CK(cudaMemset(d.error, 0, 4)); }
#ifdef B200_STC_PROFILE
    {
        unsigned long long h[8];
        CK(cudaMemcpy(h, d.prof, 64, cudaMemcpyDeviceToHost));
        printf("  [flags %u] cycles of CTA 0 warp 0: slot %llu | wait D %llu | stage_begin %llu | ld D %llu | frame_a %llu | stage_end %llu | frame_b %llu | unit end %llu\n", flags, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    }
#endif
    {
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_spectrum_tc<false>, B200_STC_THREADS, B200_STC_SMEM_BYTES));
        if (!flags) printf("  occupancy: %d CTAs / SM\n", nb);
    }
    if (!flags) printf("  tc plan: frames %u tiles/unit %u units/capture %u grid %u\n", pl.frames, pl.tiles_per_unit, pl.units_per_capture, pl.grid);
    return best;
}

int main(int argc, char **argv)
{
    const uint32_t n_cap = argc > 1 ? (uint32_t)atoi(argv[1]) : 64;
    const uint64_t big_len = argc > 2 ? (uint64_t)atoll(argv[2]) : 48000000ull;
    setvbuf(stdout, nullptr, _IONBF, 0);
    Dev d{};
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    d.sm = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, d.sm);
    CK(cudaFuncSetAttribute(k_spectrum<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B200_SPEC_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_spectrum_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B200_STC_SMEM_BYTES));

    std::vector<float> win(1024);
    for (int i = 0; i < 1024; ++i) win[i] = (float)(0.5 - 0.5 * std::cos(2.0 * b200::kPi * i / 1024.0));
    std::vector<float2> tw(1024);
    b200::fill_twiddles(tw.data());
    std::vector<uint8_t> image(B200_STC_B_BYTES);
    b200::fill_spectrum_tc_image(image.data());
    CK(cudaMalloc(&d.window, 4096)); CK(cudaMemcpy(d.window, win.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d.twiddle, 8192)); CK(cudaMemcpy(d.twiddle, tw.data(), 8192, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d.image, B200_STC_B_BYTES)); CK(cudaMemcpy(d.image, image.data(), B200_STC_B_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d.counter, 8)); CK(cudaMemset(d.counter, 0, 8));
    CK(cudaMalloc(&d.prof, 64));
    CK(cudaMalloc(&d.error, 4)); CK(cudaMemset(d.error, 0, 4));
    CK(cudaMalloc(&d.dbg, 128 * 64 * 4)); CK(cudaMemset(d.dbg, 0, 128 * 64 * 4));
    d.partial_floats = (size_t)1024 * 400 * (n_cap > 8 ? n_cap : 8);
    CK(cudaMalloc(&d.partials, d.partial_floats * 4));
    CK(cudaMalloc(&d.out32, (size_t)(n_cap > 8 ? n_cap : 8) * 4096));
    CK(cudaMalloc(&d.outtc, (size_t)(n_cap > 8 ? n_cap : 8) * 4096));

    /* ---- (1) + (2): short captures against float64 ---- */
    const uint64_t small_lens[3] = {2048 + 1024 * 14, 262144, 2 * 262144 + 1024 * 5};
    for (int t = 0; t < 3; ++t) {
        const uint64_t len = small_lens[t];
        const uint32_t nc = 3;
        std::vector<uint8_t> h(len * nc);
        for (uint32_t c = 0; c < nc; ++c) {
            std::vector<uint8_t> one(len);
            synth(one, 7 + c + 10 * t);
            std::copy(one.begin(), one.end(), h.begin() + c * len);
        }
        uint8_t *iq;
        CK(cudaMalloc(&iq, len * nc));
        CK(cudaMemcpy(iq, h.data(), len * nc, cudaMemcpyHostToDevice));
        run_fp32(d, iq, nc, len, 1);
        run_tc(d, iq, nc, len, 1, t == 0);
        std::vector<float> o32(1024 * nc), otc(1024 * nc);
        CK(cudaMemcpy(o32.data(), d.out32, 4096 * nc, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(otc.data(), d.outtc, 4096 * nc, cudaMemcpyDeviceToHost));
        const uint32_t frames = (uint32_t)b200::spectrum_frames(len);
        for (uint32_t c = 0; c < nc; ++c) {
            std::vector<double> ref(1024, 0.0);
            for (uint32_t m = 0; m < frames; ++m) {
                std::vector<std::complex<double>> x(1024);
                for (int i = 0; i < 1024; ++i) {
                    const uint8_t *s = h.data() + c * len + 2 * ((size_t)m * 512 + i);
                    const double w = 0.5 - 0.5 * std::cos(2.0 * b200::kPi * i / 1024.0);
                    x[i] = std::complex<double>((s[0] - 127.5) * w, (s[1] - 127.5) * w);
                }
                ref_fft(x);
                for (int k = 0; k < 1024; ++k) ref[k] += std::norm(x[k]);
            }
            double e32 = 0, etc = 0, pmax = 0, b32 = 0, btc = 0;
            for (int k = 0; k < 1024; ++k) { ref[k] /= frames; if (ref[k] > pmax) pmax = ref[k]; }
            for (int k = 0; k < 1024; ++k) {
                const double r32 = std::fabs(o32[c * 1024 + k] - ref[k]), rtc = std::fabs(otc[c * 1024 + k] - ref[k]);
                e32 = std::fmax(e32, r32 / ref[k]); etc = std::fmax(etc, rtc / ref[k]);
                const double bound = 1e-5 * ref[k] + 4e-7 * std::sqrt(ref[k] * pmax);
                b32 = std::fmax(b32, r32 / bound); btc = std::fmax(btc, rtc / bound);
            }
            if (t == 1 && c == 0) {
                double byc[32] = {0}, byd[32] = {0}, mean = 0;
                for (int k = 0; k < 1024; ++k) { const double e = (otc[k] - ref[k]) / ref[k]; mean += e / 1024; byc[k & 31] += e / 32; byd[k >> 5] += e / 32; }
                printf("signed rel err of tc: mean %.3e; by c:", mean);
                for (int i = 0; i < 32; ++i) printf(" %.1e", byc[i]);
                printf("\n by d:");
                for (int i = 0; i < 32; ++i) printf(" %.1e", byd[i]);
                printf("\n");
            }
            printf("len %8llu capture %u frames %5u: max rel err  fp32 %.3e  tc %.3e   (worst / few-frames bound: %.3f  %.3f)\n",
                   (unsigned long long)len, c, frames, e32, etc, b32, btc);
        }
        if (t == 0) { /* raw accumulators of half-tile 0 of unit 0: lane (w, b) = frame 2 w, column b */
            std::vector<float> y(128 * 64);
            CK(cudaMemcpy(y.data(), d.dbg, y.size() * 4, cudaMemcpyDeviceToHost));
            double worst = 0, big = 0;
            for (int w = 0; w < 4; ++w)
                for (int b = 0; b < 32; ++b)
                    for (int c = 0; c < 32; ++c) {
                        std::complex<double> acc(0, 0);
                        for (int a = 0; a < 32; ++a) {
                            const uint8_t *s = h.data() + 2 * ((size_t)(2 * w) * 512 + 32 * a + b);
                            const double ang = -2.0 * b200::kPi * ((a * c) & 31) / 32.0;
                            acc += std::complex<double>(s[0] - 128.0, s[1] - 128.0) * std::complex<double>(std::cos(ang), std::sin(ang));
                        }
                        const double er = std::fabs(y[(w * 32 + b) * 64 + 2 * c] - acc.real()), ei = std::fabs(y[(w * 32 + b) * 64 + 2 * c + 1] - acc.imag());
                        worst = std::fmax(worst, std::fmax(er, ei));
                        big = std::fmax(big, std::abs(acc));
                    }
            printf("pass-1 accumulators vs float64 DFT-32: worst abs error %.3e (largest |Y0| %.1f -> %.2e relative)\n", worst, big, worst / big);
        }
        CK(cudaFree(iq));
    }

    /* ---- (3) timing on a resident batch ---- */
    {
        std::vector<uint8_t> one(big_len);
        synth(one, 99);
        uint8_t *iq;
        CK(cudaMalloc(&iq, big_len * n_cap));
        for (uint32_t c = 0; c < n_cap; ++c) CK(cudaMemcpy(iq + c * big_len, one.data(), big_len, c == 0 ? cudaMemcpyHostToDevice : cudaMemcpyHostToDevice));
        const double samples = (double)n_cap * big_len / 2;
        for (int round = 0; round < 2; ++round) {
            const float m32 = run_fp32(d, iq, n_cap, big_len, 3);
            const float mtc = run_tc(d, iq, n_cap, big_len, 3, false);
            printf("batch %u x %llu bytes: fp32 engine %.3f ms = %.1f GS/s (%.4f of 6545.9 GB/s)   tc engine %.3f ms = %.1f GS/s (%.4f)   ratio %.3f\n", n_cap,
                   (unsigned long long)big_len, m32, samples / m32 * 1e-6, 2 * samples / m32 * 1e-6 / 6545.9, mtc, samples / mtc * 1e-6,
                   2 * samples / mtc * 1e-6 / 6545.9, m32 / mtc);
        }
        const uint32_t fl[] = {1, 2, 4, 8, 16, 1 | 16, 2 | 4 | 8, 1 | 2 | 4 | 8 | 16, 2 | 8, 1 | 8};
        for (uint32_t f : fl) printf("  dbg_flags %2u: %.3f ms\n", f, run_tc(d, iq, n_cap, big_len, 2, false, f));
        run_tc(d, iq, n_cap, big_len, 1, false);
        std::vector<float> o32(1024), otc(1024);
        CK(cudaMemcpy(o32.data(), d.out32, 4096, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(otc.data(), d.outtc, 4096, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int k = 0; k < 1024; ++k) worst = std::fmax(worst, std::fabs(o32[k] - otc[k]) / o32[k]);
        printf("big capture: max rel difference tc vs fp32 engine %.3e\n", worst);
    }
    return 0;
}
