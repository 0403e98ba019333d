/*
 * golden.h -- CPU golden model ("oracle B") of the IQ sample-processing path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or executed by
 * the product (stm32f7-rtlsdr_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker or as the timed CPU baseline.
 *
 * Pinning status:
 *   - ingest (byte identity, word granularity, block cadence): pinned bit-exactly against the
 *     UNMODIFIED reference C compiled on the host (oracle/ref_host -> oracle/_ref/), see
 *     tests/test_oracle_ref_host.py.
 *   - DSP stages (u8->cf32, window, FFT, |X|^2, averaging, FIR decimation, FM/AM demod,
 *     resampling): the reference implements NONE of them (reference README.md:29-34 lists them
 *     as next tasks; CMSIS/core/arm_math.h is vendored but never included and no CMSIS-DSP
 *     source is present), so there is no reference output to pin against: PARITY UNPINNED BY
 *     THE REFERENCE for these stages.  Per BASELINE.json north_star they are defined by this
 *     float64 model; the model itself is cross-checked against independent numpy/scipy
 *     implementations in tests/test_oracle_golden.py and against committed vectors in
 *     tests/golden/.
 *
 * Build: `make -C oracle` -> oracle/libgolden.so (float64, parity) and oracle/libgolden_f32.so
 * (same source with -DGOLD_F32: float arithmetic, used only for CPU timing).
 */
#ifndef GOLDEN_H
#define GOLDEN_H
#include <stddef.h>
#include <stdint.h>

#ifdef GOLD_F32
typedef float real;
#else
typedef double real;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GOLD_NFFT 1024
#define GOLD_HOP 512

/* window kinds / averaging modes / tap sets: same numbering as include/b200sdr.h */
enum { GOLD_WIN_RECT = 0, GOLD_WIN_HANN = 1, GOLD_WIN_BLACKMAN = 2 };
enum { GOLD_AVG_MEAN = 0, GOLD_AVG_EMA = 1 };
enum { GOLD_TAPS_FM1 = 0, GOLD_TAPS_FM2 = 1, GOLD_TAPS_AM1 = 2, GOLD_TAPS_AM2 = 3, GOLD_TAPS_AM3 = 4 };

int gold_sizeof_real(void);

/* restatement of USB_ReadPacket (stm32f7xx_ll_usb.c:792-803); returns bytes written (whole words) */
size_t gold_ingest_copy(uint8_t *dest, const uint8_t *fifo_bytes, uint16_t len);
/* test-mode counter stream: number of breaks, *first_break = smallest index or UINT64_MAX */
uint64_t gold_counter_check(const uint8_t *u, size_t len, int expect_first, uint64_t *first_break);
const float *gold_synth_lut(void);

/* x[n] = (I - 127.5) + j (Q - 127.5); out is interleaved re,im; n complex samples */
void gold_convert(const uint8_t *iq, size_t n, real *out);
/* same, times w[n mod 1024] of the given window (K2 parity) */
void gold_convert_window(const uint8_t *iq, size_t n, int window, real *out);
void gold_window(int kind, int n, real *w);
/* in-place forward DFT, X[k] = sum x[n] exp(-2 pi i k n / N), N power of two <= 65536 */
void gold_fft(real *re, real *im, int n);
/* naive O(N^2) DFT in long double, to validate gold_fft */
void gold_dft_naive(const double *re, const double *im, int n, double *ore, double *oim);

/* averaged power spectrum of one capture; returns number of frames */
uint64_t gold_spectrum(const uint8_t *iq, size_t n, int window, int avg_mode, double beta, real *out1024);

/* Kaiser-windowed-sinc low-pass: ntaps, cutoff in cycles/sample, Kaiser beta, sum(h) = gain */
void gold_kaiser_lowpass(int ntaps, double fc, double beta, double gain, double *h);
int gold_taps(int which, double *h); /* returns ntaps; h must hold >= 256 */
double gold_deemph_alpha(void);
double gold_dcblock_rho(void);

size_t gold_wbfm_disc_len(size_t n);
size_t gold_wbfm_audio_len(size_t n);
size_t gold_am_audio_len(size_t n);
/* WBFM chain; disc may be NULL; audio must hold gold_wbfm_audio_len(n) */
void gold_wbfm(const uint8_t *iq, size_t n, real *audio, real *disc);
/* stage-1 output only (interleaved re,im at 240 kS/s), for localising parity failures */
void gold_wbfm_stage1(const uint8_t *iq, size_t n, real *y1);
void gold_am(const uint8_t *iq, size_t n, real *audio);

/* 480 x 272 ARGB8888 bar plot of a float32 power spectrum (FFT order), dB window [db_min, db_max] */
void gold_render_thresholds(double db_min, double db_max, float *thr272);
void gold_render_spectrum(const float *power1024, double db_min, double db_max, uint32_t *argb);
void gold_render_waterfall(const float *spectra, uint32_t n_rows, double db_min, double db_max, uint32_t *argb);

/* synthetic captures (include/b200sdr_synth.h), n_captures x len_each bytes */
void gold_synth_fill(uint8_t *iq, uint32_t n_captures, uint64_t len_each, uint32_t kind, uint64_t first_capture);

/* ---- CPU baseline timing helpers: n_blocks blocks of n_each complex samples (block b reads
 * distinct buffer b % n_distinct) over `threads` POSIX threads; returns wall seconds ---- */
void gold_set_ingest_hook(size_t (*fn)(uint8_t *, const uint8_t *, size_t)); /* NULL = restated copy */
double gold_time_spectrum(const uint8_t *iq, size_t n_each, uint32_t n_distinct, uint32_t n_blocks, int threads);
double gold_time_wbfm(const uint8_t *iq, size_t n_each, uint32_t n_distinct, uint32_t n_blocks, int threads);
double gold_time_am(const uint8_t *iq, size_t n_each, uint32_t n_distinct, uint32_t n_blocks, int threads);

#ifdef __cplusplus
}
#endif
#endif
