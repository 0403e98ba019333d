/*
 * b200sdr.h -- C ABI of libb200sdr: the B200-native IQ sample-processing path that
 * replaces "what happens to the sample buffer" in vpecanins/stm32f7-rtlsdr.
 *
 * Reference interfaces this boundary stands in for (paths relative to the reference tree;
 * RTL/ = Middlewares/ST/STM32_USB_Host_Library/Class/RTLSDR/, USBH/ = .../Core/):
 *
 *   - The class plug-in slot `USBH_ClassTypeDef.BgndProcess` (USBH/Inc/usbh_def.h:436-447),
 *     implemented by `USBH_RTLSDR_Process` (RTL/Src/usbh_rtlsdr.c:1058-1101): once the
 *     bulk-IN URB is done the FSM sits in RTLSDR_XFER_COMPLETE (:1094-1097) with
 *     `CommItf.buff[0 .. buffSize)` (RTL/Inc/usbh_rtlsdr.h:165-173) holding u8 interleaved
 *     I,Q,I,Q...  `process_samples()` below is the consumer the author sketched in the
 *     commented-out poll at src/main.c:76-79.
 *   - Status codes are `USBH_StatusTypeDef` (USBH/Inc/usbh_def.h:301-310) by value.
 *   - The buffer rule "all buffers must have a length that is a multiple of 4 bytes"
 *     (RTL/Inc/usbh_rtlsdr.h:256-261; word-granular FIFO copy `USB_ReadPacket`,
 *     HAL_Driver/Src/stm32f7xx_ll_usb.c:792-803) is kept: `len % 4 != 0` is NOT_SUPPORTED.
 *
 * Everything here is plain C: pointers and sizes, no C++/torch types.  All device work is
 * hand-written CUDA for sm_100a inside the library; there is no CPU fallback -- if no
 * CUDA device is usable `b200sdr_create` fails with B200SDR_FAIL.
 */
#ifndef B200SDR_H
#define B200SDR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200SDR_API __attribute__((visibility("default")))
#else
#define B200SDR_API
#endif

/* ---- status: numerically identical to USBH_StatusTypeDef (USBH/Inc/usbh_def.h:301-310) ---- */
#define B200SDR_OK                  0 /* USBH_OK */
#define B200SDR_BUSY                1 /* USBH_BUSY: ring full / result not ready -- call again */
#define B200SDR_FAIL                2 /* USBH_FAIL: CUDA error or bad handle */
#define B200SDR_NOT_SUPPORTED       3 /* USBH_NOT_SUPPORTED: bad configuration / bad length */
#define B200SDR_UNRECOVERED_ERROR   4 /* USBH_UNRECOVERED_ERROR */

/* ---- processing chains (bit mask) ---- */
#define B200SDR_CHAIN_SPECTRUM 1u /* u8 -> window -> 1024-pt FFT -> |X|^2 -> average      */
#define B200SDR_CHAIN_WBFM     2u /* u8 -> /10 polyphase FIR -> discriminator -> de-emph -> /5 -> 48 kHz */
#define B200SDR_CHAIN_AM       4u /* u8 -> /20 -> /10 -> envelope -> DC block -> x2//3 -> 8 kHz */
#define B200SDR_CHAIN_COUNTER  8u /* test-mode counter check (lost samples), see b200sdr_counter_check; streaming only */

#define B200SDR_WINDOW_RECT     0u
#define B200SDR_WINDOW_HANN     1u /* periodic: w[n] = 0.5 - 0.5 cos(2 pi n / N)           */
#define B200SDR_WINDOW_BLACKMAN 2u /* periodic: 0.42 - 0.5 cos(2 pi n/N) + 0.08 cos(4 pi n/N) */

#define B200SDR_AVG_MEAN 0u /* P[k] = (1/F) sum_m |X_m[k]|^2                              */
#define B200SDR_AVG_EMA  1u /* P <- (1-beta) P + beta |X_m|^2 per frame, P starts at 0    */

/* stage-1 FIR engine of the batched WBFM path (b200sdr_batch_wbfm_dev, b200sdr_batch_host).
 * FP32: input-partitioned FIR on the CUDA cores, packed fp32x2 FMAs, taps in registers (csrc/wbfm.cuh) -- the default and
 *       the only engine of the streaming path.
 * TENSOR: the FIR as an exact u8 x s8 -> s32 banded-Toeplitz product on the tensor cores (tcgen05.mma kind::i8, TMEM),
 *       taps as three signed 8-bit slices (csrc/wbfm_tc.cuh); same results within the tolerances of the parity tests. */
#define B200SDR_FIR_ENGINE_FP32   0u
#define B200SDR_FIR_ENGINE_TENSOR 1u

#define B200SDR_NFFT 1024u
#define B200SDR_HOP  512u

/* fixed DSP parameters (frozen in oracle/golden.c) */
#define B200SDR_FS            2400000.0
#define B200SDR_FM_DECIM1     10u
#define B200SDR_FM_TAPS1      80u
#define B200SDR_FM_DECIM2     5u
#define B200SDR_FM_TAPS2      50u
#define B200SDR_AM_DECIM1     20u
#define B200SDR_AM_TAPS1      80u
#define B200SDR_AM_DECIM2     10u
#define B200SDR_AM_TAPS2      120u
#define B200SDR_AM_UP3        2u
#define B200SDR_AM_DECIM3     3u
#define B200SDR_AM_TAPS3      48u

/* tap-set selectors for b200sdr_get_taps */
#define B200SDR_TAPS_FM1 0u
#define B200SDR_TAPS_FM2 1u
#define B200SDR_TAPS_AM1 2u
#define B200SDR_TAPS_AM2 3u
#define B200SDR_TAPS_AM3 4u

typedef struct b200sdr_ctx b200sdr_ctx;

typedef struct b200sdr_config {
    uint32_t struct_size;   /* sizeof(b200sdr_config), for ABI evolution                      */
    int32_t  device;        /* CUDA device ordinal                                            */
    uint32_t chains;        /* B200SDR_CHAIN_* mask run by process_samples()                  */
    uint32_t window;        /* B200SDR_WINDOW_*                                               */
    uint32_t avg_mode;      /* B200SDR_AVG_*                                                  */
    float    ema_beta;      /* used when avg_mode == EMA                                      */
    uint32_t ring_slots;    /* pinned host ring: number of slots (>= 2)                       */
    uint32_t slot_bytes;    /* bytes per slot = largest `len` process_samples accepts.
                               Reference live value 512 (usbh_rtlsdr.c:230); BASELINE block
                               262144 = DEFAULT_BUF_LENGTH (usbh_rtlsdr.h:277-278).            */
    uint32_t audio_capacity;/* streaming audio FIFO capacity in float samples per chain; raised by
                               b200sdr_create to what ONE full slot produces (slot_bytes/100 + 16 with
                               WBFM, slot_bytes/600 + 16 with AM) so a block always fits an empty FIFO */
    uint32_t submit_bytes;  /* process_samples() appends blocks to the open pinned slot and
                               submits it (one H2D + one pass of the chains) once this many
                               bytes are pending or the next block would not fit; 0 = slot_bytes.
                               The firmware's 512-byte URBs (usbh_rtlsdr.c:230) then cost one
                               memcpy each; 4 submits every block on its own.                 */
    uint32_t fir_engine;    /* B200SDR_FIR_ENGINE_*: how the batched WBFM path computes its first (/10, 80-tap) FIR  */
    uint32_t reserved[5];
} b200sdr_config;

/* Fill *cfg with defaults: device 0, all chains, Hann, mean, 8 slots of 262144 bytes. */
B200SDR_API void b200sdr_default_config(b200sdr_config *cfg);

/* Create/destroy: the analogue of the class Init/DeInit slots
 * (USBH_RTLSDR_InterfaceInit / InterfaceDeInit, RTL/Src/usbh_rtlsdr.c:163-260, :621-641),
 * which malloc/free the handle and set up the sample buffer descriptor. */
B200SDR_API int32_t b200sdr_create(const b200sdr_config *cfg, b200sdr_ctx **out_ctx);
B200SDR_API int32_t b200sdr_destroy(b200sdr_ctx *ctx);

/* ------------------------------------------------------------------------------------------
 * Streaming path.  THE drop-in entry point: called when a buffer is complete, i.e. where the
 * reference FSM reaches RTLSDR_XFER_COMPLETE (usbh_rtlsdr.c:1094-1097).
 *   iq  : host pointer, u8 interleaved I,Q (owned by the caller)
 *   len : bytes; must be a multiple of 4 and <= cfg.slot_bytes
 *   ctx : the b200sdr_ctx* (void* so the firmware-side prototype needs no extra header)
 * The block is appended to the open pinned ring slot (so the caller may reuse `iq` on return,
 * like the reference re-arms the same buffer at once); once cfg.submit_bytes are pending (or the
 * next block would not fit) the slot is submitted: an async H2D copy on the copy stream and the
 * enabled chains on the compute stream are enqueued.  b200sdr_sync / get_spectrum / get_audio /
 * ring_acquire submit whatever is pending first, so results always cover every accepted block.
 * Filter history, FFT overlap, discriminator / IIR state are carried in ctx from call to call,
 * so a stream cut into blocks at any 4-byte boundaries yields exactly the result of one long
 * block.  Returns B200SDR_BUSY when every ring slot is still in flight (call again; nothing of
 * the block has been taken).
 * ------------------------------------------------------------------------------------------ */
B200SDR_API int32_t process_samples(const uint8_t *iq, uint32_t len, void *ctx);

/* Zero-copy variant of the above: get the next free pinned slot to fill in place
 * (what `CommItf.buff` is to the USB core), then commit `len` bytes of it. */
B200SDR_API int32_t b200sdr_ring_acquire(b200sdr_ctx *ctx, uint8_t **slot, uint32_t *slot_bytes);
B200SDR_API int32_t b200sdr_ring_commit(b200sdr_ctx *ctx, uint32_t len);

/* Block until everything enqueued so far has finished on the device. */
B200SDR_API int32_t b200sdr_sync(b200sdr_ctx *ctx);
/* Forget all carried state (new capture). */
B200SDR_API int32_t b200sdr_reset(b200sdr_ctx *ctx);

/* Current averaged power spectrum, 1024 floats in FFT order (k = 0 is DC, k = 512 is -fs/2).
 * Units: (u8 counts)^2, no normalisation.  *n_frames (optional) = frames averaged so far. */
B200SDR_API int32_t b200sdr_get_spectrum(b200sdr_ctx *ctx, float *out1024, uint64_t *n_frames);
/* Pop up to `capacity` demodulated audio samples of one chain (WBFM: 48 kHz, AM: 8 kHz). */
B200SDR_API int32_t b200sdr_get_audio(b200sdr_ctx *ctx, uint32_t chain, float *out, uint32_t capacity,
                                      uint32_t *n_out);
/* Bytes ingested so far / ring overruns (BUSY returns) -- the ingest counters the reference
 * only prints ("Xfer complete %d B, %d kB/s", usbh_rtlsdr.c:1084). */
B200SDR_API int32_t b200sdr_get_counters(b200sdr_ctx *ctx, uint64_t *bytes_in, uint64_t *blocks_in,
                                         uint64_t *busy_returns);

/* ------------------------------------------------------------------------------------------
 * Batched path: n_captures independent captures of len_each bytes, contiguous
 * (capture c starts at iq + c*len_each), each processed from zero state.
 * `*_dev` functions take DEVICE pointers (inputs already resident in HBM: this is what the
 * roofline number is measured on); `*_host` functions take HOST pointers and stream the
 * captures through the pinned ring / copy stream, overlapping H2D with compute, and write
 * results to host memory (this is the end-to-end number).
 *   spectrum out : n_captures x 1024 floats
 *   wbfm out     : n_captures x b200sdr_wbfm_audio_len(len_each) floats (48 kHz)
 *   am out       : n_captures x b200sdr_am_audio_len(len_each) floats (8 kHz)
 *   disc_out     : optional (may be NULL): raw discriminator, radians, at 240 kS/s,
 *                  n_captures x b200sdr_wbfm_disc_len(len_each)
 * All calls are asynchronous on the ctx compute stream for _dev (use b200sdr_sync) and
 * synchronous for _host.
 * ------------------------------------------------------------------------------------------ */
B200SDR_API int32_t b200sdr_batch_spectrum_dev(b200sdr_ctx *ctx, const uint8_t *iq_dev, uint32_t n_captures,
                                               uint64_t len_each, float *spectrum_dev);
B200SDR_API int32_t b200sdr_batch_wbfm_dev(b200sdr_ctx *ctx, const uint8_t *iq_dev, uint32_t n_captures,
                                           uint64_t len_each, float *audio_dev, float *disc_dev);
B200SDR_API int32_t b200sdr_batch_am_dev(b200sdr_ctx *ctx, const uint8_t *iq_dev, uint32_t n_captures,
                                         uint64_t len_each, float *audio_dev);
B200SDR_API int32_t b200sdr_batch_host(b200sdr_ctx *ctx, uint32_t chains, const uint8_t *iq_host,
                                       uint32_t n_captures, uint64_t len_each, float *spectrum_host,
                                       float *wbfm_audio_host, float *am_audio_host);

/* ------------------------------------------------------------------------------------------
 * Optional exchange step: ONE long capture split in time across the
 * GPUs of a box for latency.  Rank r (one process / one ctx per GPU) holds the slice of the
 * capture that contains its frames (stm32f7-rtlsdr_b200/sharding.py split_capture_bytes: its
 * frames plus the 512-sample FFT overlap) and calls b200sdr_split_spectrum_dev; ONE fused
 * kernel reduces the local partial sums, pushes them into every peer's mailbox over NVLink
 * peer memory, waits for all ranks and adds the contributions in rank order, so every rank
 * ends with bitwise the same mean spectrum of the whole capture.  No library collective.
 * Setup: every rank calls exchange_create (gets a 64-byte IPC handle of its mailbox), the
 * handles are gathered by any means (bench/tests use torch.distributed as plumbing) and
 * passed in rank order to exchange_connect.  Contexts living in one process use
 * exchange_connect_local instead.  All ranks must make the same sequence of split calls.
 * split_spectrum_dev is asynchronous on the ctx compute stream; exchange_wait blocks and
 * returns B200SDR_FAIL if a peer did not arrive within ~5 s (the kernel's wait is bounded); the
 * output spectrum is then NaN-filled, and because the ranks no longer agree on the call number the
 * exchange must be destroyed and re-created (exchange_destroy / _create / _connect) on EVERY rank
 * before the next split call.
 * Demodulators are not split this way (carried IIR state): they stay one capture per GPU.
 * ------------------------------------------------------------------------------------------ */
B200SDR_API int32_t b200sdr_exchange_create(b200sdr_ctx *ctx, uint32_t world, uint32_t rank, uint8_t *ipc_handle_out64);
B200SDR_API int32_t b200sdr_exchange_connect(b200sdr_ctx *ctx, const uint8_t *ipc_handles /* world x 64 bytes */);
B200SDR_API int32_t b200sdr_exchange_connect_local(b200sdr_ctx *ctx, b200sdr_ctx *const *peers /* world ctxs */);
B200SDR_API int32_t b200sdr_split_spectrum_dev(b200sdr_ctx *ctx, const uint8_t *iq_slice_dev, uint64_t len_slice,
                                               uint64_t frames_total, float *spectrum_dev);
B200SDR_API int32_t b200sdr_exchange_wait(b200sdr_ctx *ctx);
B200SDR_API int32_t b200sdr_exchange_destroy(b200sdr_ctx *ctx);

B200SDR_API uint64_t b200sdr_spectrum_frames(uint64_t len_bytes);  /* floor((L-1024)/512)+1, 0 if L<1024 */
B200SDR_API uint64_t b200sdr_wbfm_disc_len(uint64_t len_bytes);    /* ceil(L/10)                          */
B200SDR_API uint64_t b200sdr_wbfm_audio_len(uint64_t len_bytes);   /* ceil(ceil(L/10)/5)                  */
B200SDR_API uint64_t b200sdr_am_audio_len(uint64_t len_bytes);     /* ceil(2*ceil(ceil(L/20)/10)/3)       */
/* Streaming granularity of a demodulator chain (B200SDR_CHAIN_WBFM / _AM) in complex samples: the streaming
 * path consumes whole chunks, what is left of a block (< one chunk) waits for the next block.  After n
 * accepted samples the audio FIFO has received ceil(floor(n / chunk) * (chunk / 10) / 5) WBFM samples
 * (AM: (2 floor(n / chunk) + 2) / 3).  0 for a chain without audio output. */
B200SDR_API uint32_t b200sdr_stream_chunk_samples(uint32_t chain);

/* ------------------------------------------------------------------------------------------
 * Stand-alone IQ conversion (kernel K2): out[2n] = (I_n - 127.5) * w[n mod 1024],
 * out[2n+1] = (Q_n - 127.5) * w[n mod 1024]; window RECT gives the exact, unscaled
 * conversion.  Host pointers; `len` bytes in, `len` floats out.
 * ------------------------------------------------------------------------------------------ */
B200SDR_API int32_t b200sdr_convert_cf32(b200sdr_ctx *ctx, const uint8_t *iq_host, uint32_t len, uint32_t window,
                                         float *out_host);
B200SDR_API int32_t b200sdr_convert_cf32_dev(b200sdr_ctx *ctx, const uint8_t *iq_dev, uint64_t len, uint32_t window,
                                             float *out_dev);

/* ------------------------------------------------------------------------------------------
 * Test-mode counter check (kernel K0).  The firmware puts the RTL2832 into test mode before it starts
 * streaming (RTLSDR_set_test_mode(phost, 1), RTL/Src/usbh_rtlsdr.c:901, :660-662), so the bytes that reach
 * the buffer boundary are the dongle's 8-bit counter and every byte that is not its predecessor + 1 marks
 * lost samples.  Per capture: n_breaks = number of indices i >= 1 with u[i] != (u[i-1] + 1) & 0xff, plus one
 * for i = 0 when expect_first is 0..255 and u[0] differs from it (chain blocks with expect_first = last byte
 * of the previous block + 1; -1 = do not check the first byte); first_break = the smallest such index or
 * UINT64_MAX.  Results are written to HOST arrays of n_captures entries (either may be NULL); the call returns
 * when they are valid.  `_dev`: captures resident in HBM, len_each a multiple of 4, 16-byte aligned pointer and
 * stride (len_each a multiple of 16 unless n_captures == 1).  Host variant: one block of `len` bytes.
 * ------------------------------------------------------------------------------------------ */
B200SDR_API int32_t b200sdr_counter_check_dev(b200sdr_ctx *ctx, const uint8_t *iq_dev, uint32_t n_captures, uint64_t len_each,
                                              int32_t expect_first, uint64_t *n_breaks_host, uint64_t *first_break_host);
B200SDR_API int32_t b200sdr_counter_check(b200sdr_ctx *ctx, const uint8_t *iq_host, uint32_t len, int32_t expect_first,
                                          uint64_t *n_breaks, uint64_t *first_break);
/* Streaming: with B200SDR_CHAIN_COUNTER in cfg.chains every block accepted by process_samples() is checked as a
 * continuation of the previous one (the last byte is carried on the device).  Totals since create / reset:
 * *n_breaks, and *first_break = index of the first offending byte counted from the start of the stream
 * (UINT64_MAX: none).  Covers every accepted block (pending bytes are submitted first, like b200sdr_get_spectrum). */
B200SDR_API int32_t b200sdr_get_counter_check(b200sdr_ctx *ctx, uint64_t *n_breaks, uint64_t *first_break);

/* ------------------------------------------------------------------------------------------
 * Presentation: a power spectrum as a 480 x 272 ARGB8888 bar plot --
 * the geometry and pixel format of the LCD layer the firmware's sample buffer aliases
 * (src/main.c:100-109).  Column 0 is -fs/2, the centre column is DC; dB = 10 log10(power) mapped
 * from [db_min, db_max] to the 272 rows.  spectrum_host == NULL renders the current streaming
 * spectrum.  `_dev` renders n spectra resident on the device (values are multiplied by `scale`).
 * ------------------------------------------------------------------------------------------ */
B200SDR_API int32_t b200sdr_render_spectrum(b200sdr_ctx *ctx, const float *spectrum_host, float db_min, float db_max,
                                            uint32_t *argb_host /* 480*272 */);
B200SDR_API int32_t b200sdr_render_spectrum_dev(b200sdr_ctx *ctx, const float *spectra_dev, uint32_t n_spectra,
                                                float scale, float db_min, float db_max, uint32_t *argb_dev);
/* Waterfall (spectrogram) view on the same panel: image row r shows spectrum r of `n_rows` x 1024
 * (e.g. the output of b200sdr_batch_spectrum_dev over consecutive slices of a capture) in the colour
 * of the bar height that power would reach; black below db_min and for rows >= n_rows; at most the
 * first 272 spectra are shown. */
B200SDR_API int32_t b200sdr_render_waterfall(b200sdr_ctx *ctx, const float *spectra_host, uint32_t n_rows, float db_min,
                                             float db_max, uint32_t *argb_host /* 480*272 */);
B200SDR_API int32_t b200sdr_render_waterfall_dev(b200sdr_ctx *ctx, const float *spectra_dev, uint32_t n_rows,
                                                 float scale, float db_min, float db_max, uint32_t *argb_dev);

/* The FIR taps / window the device uses (float, as uploaded), for parity against the oracle. */
B200SDR_API int32_t b200sdr_get_taps(b200sdr_ctx *ctx, uint32_t which, float *out, uint32_t capacity,
                                     uint32_t *n_taps);
B200SDR_API int32_t b200sdr_get_window(b200sdr_ctx *ctx, uint32_t window, float *out1024);

/* Parity hook of the TENSOR FIR engine: run it over ONE device-resident capture (len % 16 == 0) and return the raw
 * 32-bit accumulators of its first tile, acc_host[row r][34 s + 2 j + c] = sum_t q_s[t] u[2 (10 m - t) + c] with
 * m = 16 r - 1 + j (128 x 112: slice s = 0..2, j = 0..16 -- the output before the row, then its sixteen --, component c;
 * rows 125..127 and columns 102..111 unused; u[n < 0] = 0), the three signed 8-bit tap slices (slices_host[3][80], may be
 * NULL) and the
 * exponent e with h[t] 2^e = q0 2^-7 + q1 2^-14 + q2 2^-21 (may be NULL): the integer product must be bit-exact. */
B200SDR_API int32_t b200sdr_debug_wbfm_tc_acc(b200sdr_ctx *ctx, const uint8_t *iq_dev, uint64_t len, int32_t *acc_host,
                                              int8_t *slices_host, int32_t *exponent);

/* Read back the device copy of the most recently ingested block (ingest parity checks). */
B200SDR_API int32_t b200sdr_debug_last_block(b200sdr_ctx *ctx, uint8_t *out, uint32_t capacity, uint32_t *len);

/* ------------------------------------------------------------------------------------------
 * Synthetic captures generated on the device (include/b200sdr_synth.h), for benchmarks:
 * capture c gets seed B200SDR_SYNTH_SEED_BASE + first_capture + c.
 * ------------------------------------------------------------------------------------------ */
B200SDR_API int32_t b200sdr_synth_fill_dev(b200sdr_ctx *ctx, uint8_t *iq_dev, uint32_t n_captures, uint64_t len_each,
                                           uint32_t kind, uint64_t first_capture);
/* Same generator on the host (no device work), so callers can feed the streaming path. */
B200SDR_API int32_t b200sdr_synth_fill_host(uint8_t *iq_host, uint32_t n_captures, uint64_t len_each, uint32_t kind,
                                            uint64_t first_capture);

/* Device memory helpers so a pure-C host driver needs no CUDA headers. */
B200SDR_API int32_t b200sdr_dev_alloc(b200sdr_ctx *ctx, uint64_t bytes, void **out_dev);
B200SDR_API int32_t b200sdr_dev_free(b200sdr_ctx *ctx, void *dev);
B200SDR_API int32_t b200sdr_host_alloc_pinned(b200sdr_ctx *ctx, uint64_t bytes, void **out_host);
B200SDR_API int32_t b200sdr_host_free_pinned(b200sdr_ctx *ctx, void *host);
B200SDR_API int32_t b200sdr_copy_to_host(b200sdr_ctx *ctx, void *dst_host, const void *src_dev, uint64_t bytes);
B200SDR_API int32_t b200sdr_copy_to_dev(b200sdr_ctx *ctx, void *dst_dev, const void *src_host, uint64_t bytes);

/* Timing of device work on the ctx compute stream with CUDA events (bench.py). */
B200SDR_API int32_t b200sdr_timer_start(b200sdr_ctx *ctx);
B200SDR_API int32_t b200sdr_timer_stop_ms(b200sdr_ctx *ctx, float *ms);
/* Number of kernels this library has launched since ctx creation. */
B200SDR_API uint64_t b200sdr_kernel_launches(b200sdr_ctx *ctx);
B200SDR_API const char *b200sdr_last_error(b200sdr_ctx *ctx);
B200SDR_API const char *b200sdr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SDR_H */
