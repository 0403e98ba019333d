/*
 * firmware_cadence.c -- the reference firmware's superloop, verbatim in shape, with libb200sdr
 * dropped in at the buffer-complete point.
 *
 * The firmware owns ONE sample buffer (CommItf.buff, 512 bytes, RTL/Src/usbh_rtlsdr.c:227-230);
 * USBH_RTLSDR_Process (:1058-1101) submits a URB into it, waits, and reaches RTLSDR_XFER_COMPLETE
 * with the buffer full -- then immediately re-arms the SAME buffer.  Here the "USB core" is a memcpy
 * from a synthetic stream into that one buffer, and COMPLETE calls process_samples(buff, 512, ctx).
 * Every `refresh` blocks the loop reads the spectrum back, like an LCD refresh would.
 *
 *   gcc -O2 -Iinclude examples/firmware_cadence.c -Lstm32f7-rtlsdr_b200 -lb200sdr \
 *       -Wl,-rpath,'$ORIGIN/../stm32f7-rtlsdr_b200' -o build/firmware_cadence
 *   build/firmware_cadence [total_bytes] [block_bytes] [submit_bytes] [slot_bytes] [refresh_blocks]
 *
 * Prints one line:  FW_CADENCE block=.. submit=.. blocks=.. busy=.. seconds=.. MSps=.. us_per_block=.. realtime=..
 * (realtime = throughput / the dongle's 2.4 MS/s).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200sdr.h"
#include "b200sdr_synth.h"

enum { XFER_START = 0, XFER_WAIT, XFER_COMPLETE }; /* RTLSDR_xferStateTypeDef, usbh_rtlsdr.h:156-162 */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    size_t total = argc > 1 ? strtoull(argv[1], 0, 0) : (size_t)48000000;
    uint32_t block = argc > 2 ? (uint32_t)strtoul(argv[2], 0, 0) : 512u;
    uint32_t submit = argc > 3 ? (uint32_t)strtoul(argv[3], 0, 0) : 0u;
    uint32_t slot = argc > 4 ? (uint32_t)strtoul(argv[4], 0, 0) : 262144u;
    unsigned long refresh = argc > 5 ? strtoul(argv[5], 0, 0) : 0ul;
    block &= ~3u; /* the multiple-of-4 buffer rule, usbh_rtlsdr.h:256-261 */
    if (block == 0 || block > slot) { fprintf(stderr, "block must be 4..slot_bytes\n"); return 2; }
    total -= total % block;

    const size_t stream_bytes = (size_t)16 << 20; /* the stream repeats every 16 MiB */
    uint8_t *stream = (uint8_t *)malloc(stream_bytes + block);
    if (!stream || b200sdr_synth_fill_host(stream, 1, stream_bytes + block - (stream_bytes + block) % 4, B200SDR_SYNTH_WBFM, 0) != B200SDR_OK) return 3;

    b200sdr_config cfg;
    b200sdr_default_config(&cfg);
    cfg.chains = B200SDR_CHAIN_SPECTRUM | B200SDR_CHAIN_WBFM;
    cfg.slot_bytes = slot;
    cfg.submit_bytes = submit;
    cfg.audio_capacity = 1u << 22;
    b200sdr_ctx *ctx = NULL;
    if (b200sdr_create(&cfg, &ctx) != B200SDR_OK) { fprintf(stderr, "b200sdr_create failed: a CUDA sm_100 device is required\n"); return 4; }

    uint8_t *buff = (uint8_t *)malloc(block); /* CommItf.buff */
    float *audio = (float *)malloc(sizeof(float) * (1u << 22));
    float spec[1024];
    uint64_t frames = 0;
    unsigned long blocks = 0, busy = 0;
    size_t off = 0, pos = 0;
    int state = XFER_START;
    const double t0 = now_s();
    while (off < total) {
        switch (state) {
        case XFER_START: /* USBH_BulkReceiveData(phost, buff, buffSize, pipe) */
            state = XFER_WAIT;
            break;
        case XFER_WAIT: /* URB done: the IRQ handler has copied the packets into buff */
            memcpy(buff, stream + pos, block);
            pos += block;
            if (pos >= stream_bytes) pos = 0;
            state = XFER_COMPLETE;
            break;
        case XFER_COMPLETE: {
            const int32_t rc = process_samples(buff, block, ctx);
            if (rc == B200SDR_BUSY) { /* ring or audio FIFO full: drain, stay in COMPLETE (USBH_BUSY) */
                uint32_t n = 0;
                busy++;
                b200sdr_get_audio(ctx, B200SDR_CHAIN_WBFM, audio, 1u << 22, &n);
                break;
            }
            if (rc != B200SDR_OK) { fprintf(stderr, "process_samples: %s\n", b200sdr_last_error(ctx)); return 5; }
            off += block;
            blocks++;
            if (refresh && blocks % refresh == 0) b200sdr_get_spectrum(ctx, spec, &frames);
            state = XFER_START;
            break;
        }
        }
    }
    b200sdr_sync(ctx);
    const double dt = now_s() - t0;
    b200sdr_get_spectrum(ctx, spec, &frames);
    const double msps = (double)total / 2.0 / dt / 1e6;
    printf("FW_CADENCE block=%u submit=%u slot=%u blocks=%lu busy=%lu frames=%llu seconds=%.4f MSps=%.1f us_per_block=%.3f realtime=%.1f\n",
           block, submit ? submit : slot, slot, blocks, busy, (unsigned long long)frames, dt, msps, dt / (double)blocks * 1e6, msps / 2.4);
    free(buff); free(audio); free(stream);
    b200sdr_destroy(ctx);
    return 0;
}
