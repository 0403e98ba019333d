/*
 * selftest.c -- runs the golden model's entry points on small inputs under AddressSanitizer /
 * UndefinedBehaviorSanitizer (built and run by tests/test_oracle_sanitizers.py).
 * TEST INFRASTRUCTURE ONLY.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "golden.h"

int main(void)
{
    const size_t n = 4096 + 77;             /* ragged on purpose */
    uint8_t *iq = (uint8_t *)malloc(2 * n); /* exact size: any over-read is caught */
    double taps[256];
    for (int kind = 0; kind < 4; ++kind) gold_synth_fill(iq, 1, 2 * n, (uint32_t)kind, 3);
    real *cf = (real *)malloc(sizeof(real) * 2 * n);
    gold_convert(iq, n, cf);
    gold_convert_window(iq, n, GOLD_WIN_HANN, cf);
    real spec[1024];
    uint64_t frames = gold_spectrum(iq, n, GOLD_WIN_HANN, GOLD_AVG_MEAN, 0.0, spec);
    frames += gold_spectrum(iq, n, GOLD_WIN_BLACKMAN, GOLD_AVG_EMA, 0.1, spec);
    frames += gold_spectrum(iq, 1000, GOLD_WIN_HANN, GOLD_AVG_MEAN, 0.0, spec); /* too short: 0 frames */
    for (int t = 0; t < 5; ++t) gold_taps(t, taps);
    real *audio = (real *)malloc(sizeof(real) * (gold_wbfm_audio_len(n) + 1));
    real *disc = (real *)malloc(sizeof(real) * (gold_wbfm_disc_len(n) + 1));
    gold_wbfm(iq, n, audio, disc);
    gold_wbfm(iq, 1, audio, disc); /* a single sample */
    real *am = (real *)malloc(sizeof(real) * (gold_am_audio_len(n) + 1));
    gold_am(iq, n, am);
    gold_am(iq, 1, am);
    uint8_t dst[516];
    memset(dst, 0, sizeof dst);
    size_t w = gold_ingest_copy(dst, iq, 510);
    float power[1024];
    for (int k = 0; k < 1024; ++k) power[k] = (float)spec[k];
    uint32_t *img = (uint32_t *)malloc(sizeof(uint32_t) * 480 * 272);
    gold_render_spectrum(power, 0.0, 100.0, img);
    gold_render_waterfall(power, 1, 0.0, 100.0, img);
    gold_render_waterfall(power, 0, 0.0, 100.0, img);
    gold_render_spectrum(power, 0.0, 100.0, img);
    printf("SELFTEST_OK frames=%llu copied=%zu px=%08x\n", (unsigned long long)frames, w, img[271 * 480 + 240]);
    free(iq); free(cf); free(audio); free(disc); free(am); free(img);
    return 0;
}
