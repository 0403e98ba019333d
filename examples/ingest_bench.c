/*
 * ingest_bench.c -- what one ring slot costs the HOST, from a plain C caller of the ABI (include/b200sdr.h):
 *
 *   copy      process_samples(block, len, ctx): the block is copied into the pinned slot (the caller gets its buffer back
 *             at once, like the firmware's CommItf.buff that the next URB overwrites, RTL/Src/usbh_rtlsdr.c:1070-1097),
 *             then H2D + chains are enqueued;
 *   zero-copy b200sdr_ring_acquire / b200sdr_ring_commit: the producer (a USB bulk read, read(2), recv(2) ...) fills the
 *             pinned slot itself, so the per-slot host cost is the enqueue alone.  The producer is not part of the
 *             timed loop here (the slots are filled once, before it).
 *
 *   gcc -O2 -Iinclude examples/ingest_bench.c -Lstm32f7-rtlsdr_b200 -lb200sdr -Wl,-rpath,'$ORIGIN/../stm32f7-rtlsdr_b200' -o build/ingest_bench
 *   build/ingest_bench [block_bytes] [blocks] [chains]
 *
 * Prints:  INGEST mode=copy|zero_copy block=.. blocks=.. busy=.. us_per_block=.. GBps=.. MSps=.. realtime=.. frames=..
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200sdr.h"
#include "b200sdr_synth.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    uint32_t block = argc > 1 ? (uint32_t)strtoul(argv[1], 0, 0) : 262144u;
    unsigned long blocks = argc > 2 ? strtoul(argv[2], 0, 0) : 2048ul;
    uint32_t chains = argc > 3 ? (uint32_t)strtoul(argv[3], 0, 0) : (B200SDR_CHAIN_SPECTRUM | B200SDR_CHAIN_WBFM);
    block &= ~3u;
    if (block == 0) return 2;

    uint8_t *src = (uint8_t *)malloc((size_t)16 * block);
    if (!src || b200sdr_synth_fill_host(src, 1, (uint64_t)16 * block, B200SDR_SYNTH_WBFM, 0) != B200SDR_OK) return 3;
    float *audio = (float *)malloc(sizeof(float) * (1u << 22));
    float spec[1024];

    for (int mode = 0; mode < 2; ++mode) {
        b200sdr_config cfg;
        b200sdr_default_config(&cfg);
        cfg.chains = chains;
        cfg.slot_bytes = block;
        cfg.ring_slots = 8;
        cfg.audio_capacity = 1u << 22;
        b200sdr_ctx *ctx = NULL;
        if (b200sdr_create(&cfg, &ctx) != B200SDR_OK) { fprintf(stderr, "b200sdr_create failed: a CUDA sm_100 device is required\n"); return 4; }
        unsigned long busy = 0, filled = 0;
        double t0 = 0.0;
        for (unsigned long i = 0; i < blocks + 64; ++i) {
            if (i == 64) { /* warm-up done: drain, start the clock */
                uint32_t n = 0;
                b200sdr_sync(ctx);
                b200sdr_get_audio(ctx, B200SDR_CHAIN_WBFM, audio, 1u << 22, &n);
                t0 = now_s();
            }
            if ((i & 255ul) == 255ul) { /* keep the audio FIFO from filling up on long runs */
                uint32_t n = 0;
                if (chains & B200SDR_CHAIN_WBFM) b200sdr_get_audio(ctx, B200SDR_CHAIN_WBFM, audio, 1u << 22, &n);
                if (chains & B200SDR_CHAIN_AM) b200sdr_get_audio(ctx, B200SDR_CHAIN_AM, audio, 1u << 22, &n);
            }
            int32_t rc;
            if (mode == 0) {
                while ((rc = process_samples(src + (size_t)(i % 16) * block, block, ctx)) == B200SDR_BUSY) ++busy;
            } else {
                uint8_t *slot = NULL;
                uint32_t cap = 0;
                while ((rc = b200sdr_ring_acquire(ctx, &slot, &cap)) == B200SDR_BUSY) ++busy;
                if (rc == B200SDR_OK) {
                    if (filled < cfg.ring_slots) { memcpy(slot, src + (size_t)(i % 16) * block, block); ++filled; } /* the "producer", once per slot */
                    rc = b200sdr_ring_commit(ctx, block);
                }
            }
            if (rc != B200SDR_OK) { fprintf(stderr, "ingest failed: %s\n", b200sdr_last_error(ctx)); return 5; }
        }
        b200sdr_sync(ctx);
        const double dt = now_s() - t0;
        uint64_t frames = 0;
        b200sdr_get_spectrum(ctx, spec, &frames);
        printf("INGEST mode=%s block=%u blocks=%lu busy=%lu us_per_block=%.3f GBps=%.3f MSps=%.1f realtime=%.1f frames=%llu\n",
               mode == 0 ? "copy" : "zero_copy", block, blocks, busy, dt / (double)blocks * 1e6, (double)blocks * block / dt / 1e9,
               (double)blocks * block / 2.0 / dt / 1e6, (double)blocks * block / 2.0 / dt / 2.4e6, (unsigned long long)frames);
        b200sdr_destroy(ctx);
    }
    free(src);
    free(audio);
    return 0;
}
