/*
 * host_driver.c -- a host-side C driver that feeds libb200sdr with the block cadence of the
 * reference firmware.  The pinned ring stands in for the USB-OTG ingest: where the firmware's
 * class FSM (RTL/Src/usbh_rtlsdr.c:1058-1101) goes START -> WAIT -> COMPLETE around one URB into
 * CommItf.buff, this driver goes acquire-slot -> fill -> commit.
 *
 *   gcc -O2 -Iinclude examples/host_driver.c -Lstm32f7-rtlsdr_b200 -lb200sdr \
 *       -Wl,-rpath,'$ORIGIN/../stm32f7-rtlsdr_b200' -o build/host_driver
 *   build/host_driver capture.cu8 [block_bytes]        (rtl_sdr / GQRX .cu8 file)
 *   build/host_driver --synthetic wbfm 4800000         (bytes of the built-in FM test signal)
 *
 * Prints the strongest spectrum bins and writes fm48k.f32 / am8k.f32 next to the working dir.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200sdr.h"
#include "b200sdr_synth.h"

enum { XFER_START = 0, XFER_WAIT, XFER_COMPLETE }; /* RTLSDR_xferStateTypeDef, usbh_rtlsdr.h:156-162 */

static int drain(b200sdr_ctx *ctx, uint32_t chain, FILE *f, float *tmp, uint32_t cap)
{
    uint32_t n = 0;
    if (b200sdr_get_audio(ctx, chain, tmp, cap, &n) != B200SDR_OK) return -1;
    if (f && n) fwrite(tmp, sizeof(float), n, f);
    return (int)n;
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s file.cu8 [block_bytes] | --synthetic multitone|wbfm|am nbytes [block_bytes]\n", argv[0]);
        return 2;
    }
    uint8_t *data = NULL;
    size_t total = 0;
    uint32_t block = 262144; /* DEFAULT_BUF_LENGTH, usbh_rtlsdr.h:277-278 */
    if (strcmp(argv[1], "--synthetic") == 0 && argc >= 4) {
        uint32_t kind = strcmp(argv[2], "wbfm") == 0 ? B200SDR_SYNTH_WBFM : strcmp(argv[2], "am") == 0 ? B200SDR_SYNTH_AM : B200SDR_SYNTH_MULTITONE;
        total = strtoull(argv[3], 0, 0) & ~(size_t)3;
        data = (uint8_t *)malloc(total);
        if (b200sdr_synth_fill_host(data, 1, total, kind, 0) != B200SDR_OK) return 3;
        if (argc >= 5) block = (uint32_t)strtoul(argv[4], 0, 0);
    } else {
        FILE *f = fopen(argv[1], "rb");
        if (!f) { perror(argv[1]); return 3; }
        fseek(f, 0, SEEK_END);
        total = (size_t)ftell(f) & ~(size_t)3;
        fseek(f, 0, SEEK_SET);
        data = (uint8_t *)malloc(total);
        if (fread(data, 1, total, f) != total) return 3;
        fclose(f);
        if (argc >= 3) block = (uint32_t)strtoul(argv[2], 0, 0);
    }
    block &= ~3u; /* the multiple-of-4 buffer rule, usbh_rtlsdr.h:256-261 */

    b200sdr_config cfg;
    b200sdr_default_config(&cfg);
    cfg.slot_bytes = block;
    cfg.ring_slots = 8;
    b200sdr_ctx *ctx = NULL;
    int rc = b200sdr_create(&cfg, &ctx);
    if (rc != B200SDR_OK) { fprintf(stderr, "b200sdr_create failed (%d): a CUDA sm_100 device is required\n", rc); return 4; }

    FILE *ffm = fopen("fm48k.f32", "wb"), *fam = fopen("am8k.f32", "wb");
    float *tmp = (float *)malloc(sizeof(float) * (1u << 20));
    size_t off = 0;
    unsigned long blocks = 0, busy = 0;
    int state = XFER_START;
    uint8_t *slot = NULL;
    uint32_t slot_bytes = 0, n = 0;
    while (off < total || state != XFER_START) {
        switch (state) {
        case XFER_START: /* reference: USBH_BulkReceiveData(buff, buffSize) */
            rc = b200sdr_ring_acquire(ctx, &slot, &slot_bytes);
            if (rc == B200SDR_BUSY) { /* every slot still in flight: poll again, like USBH_BUSY */
                busy++;
                drain(ctx, B200SDR_CHAIN_WBFM, ffm, tmp, 1u << 20);
                drain(ctx, B200SDR_CHAIN_AM, fam, tmp, 1u << 20);
                break;
            }
            if (rc != B200SDR_OK) { fprintf(stderr, "acquire: %s\n", b200sdr_last_error(ctx)); return 5; }
            state = XFER_WAIT;
            break;
        case XFER_WAIT: /* reference: poll URB state; here the "USB core" is a memcpy from the file */
            n = (uint32_t)((total - off < slot_bytes) ? total - off : slot_bytes);
            memcpy(slot, data + off, n);
            off += n;
            state = XFER_COMPLETE;
            break;
        case XFER_COMPLETE: /* reference: nothing happens here -- this is the hook */
            rc = b200sdr_ring_commit(ctx, n);
            if (rc == B200SDR_BUSY) { /* audio FIFO full: drain and retry the commit */
                drain(ctx, B200SDR_CHAIN_WBFM, ffm, tmp, 1u << 20);
                drain(ctx, B200SDR_CHAIN_AM, fam, tmp, 1u << 20);
                break;
            }
            if (rc != B200SDR_OK) { fprintf(stderr, "commit: %s\n", b200sdr_last_error(ctx)); return 5; }
            blocks++;
            state = XFER_START;
            break;
        }
    }
    b200sdr_sync(ctx);
    drain(ctx, B200SDR_CHAIN_WBFM, ffm, tmp, 1u << 20);
    drain(ctx, B200SDR_CHAIN_AM, fam, tmp, 1u << 20);
    float spec[1024];
    uint64_t frames = 0, bytes_in = 0, blocks_in = 0, busy_ret = 0;
    b200sdr_get_spectrum(ctx, spec, &frames);
    b200sdr_get_counters(ctx, &bytes_in, &blocks_in, &busy_ret);
    printf("%lu blocks of <= %u bytes, %llu bytes in, %llu frames averaged, %lu busy polls\n", blocks, block,
           (unsigned long long)bytes_in, (unsigned long long)frames, busy + (unsigned long)busy_ret);
    for (int top = 0; top < 5; ++top) {
        int best = 0;
        for (int k = 1; k < 1024; ++k) if (spec[k] > spec[best]) best = k;
        double f_hz = (best < 512 ? best : best - 1024) * B200SDR_FS / 1024.0;
        printf("  bin %4d  %+10.1f Hz  power %.4g\n", best, f_hz, spec[best]);
        spec[best] = 0;
    }
    fclose(ffm); fclose(fam);
    free(tmp); free(data);
    b200sdr_destroy(ctx);
    return 0;
}
