/*
 * emu_kernels.cpp -- runs the product's kernel SOURCE (stm32f7-rtlsdr_b200/csrc/*.cuh) on the
 * host under tests/emu/cuda_emu.h, so the CPU test-suite can check kernel index arithmetic
 * against the oracle without a GPU.  TEST INFRASTRUCTURE ONLY -- not a product fallback.
 * Built by tests/conftest.py:  g++ -O1 -shared -fPIC -o tests/emu/libemu_kernels.so
 */
#include "cuda_emu.h"

#include <cstring>
#include <vector>

#include "../../stm32f7-rtlsdr_b200/csrc/plan.h"
#include "../../stm32f7-rtlsdr_b200/csrc/misc_kernels.cuh"
#include "../../stm32f7-rtlsdr_b200/csrc/render.cuh"

extern "C" {

/* mean / EMA spectrum of one or more captures through k_spectrum + k_spectrum_finalize */
int emu_spectrum(const uint8_t *iq, uint32_t n_captures, uint64_t len_each_bytes, const float *window,
                 uint32_t frames_per_warp, int ema, float beta, float *out)
{
    uint32_t frames = (uint32_t)b200::spectrum_frames(len_each_bytes);
    if (frames == 0) return -1;
    std::vector<float2> tw(1024);
    b200::fill_twiddles(tw.data());
    uint32_t frames_per_cta = frames_per_warp * B200_SPEC_WARPS;
    uint32_t ctas = (frames + frames_per_cta - 1) / frames_per_cta;
    std::vector<float> partials((size_t)n_captures * ctas * 1024);
    SpectrumParams p;
    p.iq = iq;
    p.capture_stride = len_each_bytes;
    p.frames = frames;
    p.frames_per_warp = frames_per_warp;
    p.window = window;
    p.twiddle = tw.data();
    p.partials = partials.data();
    p.units_per_capture = ctas;
    p.total_units = ctas * n_captures;
    uint32_t unit_counter[2] = {0u, 0u};
    p.unit_counter = unit_counter;
    p.final_out = nullptr;
    p.carry = nullptr;
    p.final_scale = p.carry_scale = 0.0f;
    std::vector<float> folded(1024, 0.0f);
    if (n_captures == 1) { /* the streaming form: the last CTA finalizes; must equal k_spectrum_finalize bit for bit */
        p.final_out = folded.data();
        p.final_scale = ema ? 1.0f : 1.0f / (float)frames;
    }
    p.ema_beta = beta;
    p.ema_log2_decay = log2f(1.0f - beta);
    /* a persistent grid smaller than the number of units: every CTA walks several units */
    const uint32_t grid = p.total_units > 3 ? 3 : p.total_units;
    if (ema) emu::launch(dim3(grid), dim3(B200_SPEC_THREADS), B200_SPEC_SMEM_BYTES, [&] { k_spectrum<true>(p); });
    else emu::launch(dim3(grid), dim3(B200_SPEC_THREADS), B200_SPEC_SMEM_BYTES, [&] { k_spectrum<false>(p); });
    float scale = ema ? 1.0f : 1.0f / (float)frames;
    if (unit_counter[0] != 0u || unit_counter[1] != 0u) return -2; /* the last CTA must leave the hand-out counter at zero */
    emu::launch(dim3(4, n_captures), dim3(256), 0,
                [&] { k_spectrum_finalize(partials.data(), ctas, scale, nullptr, 0.0f, out); });
    if (n_captures == 1 && memcmp(folded.data(), out, 1024 * sizeof(float)) != 0) return -3;
    return (int)frames;
}

/* WBFM over whole captures; tiles_per_segment = 0 takes the product's plan.
 * `iq` must be readable up to a multiple of 16 bytes (capture_bytes is rounded down to 16 like the
 * batched API requires). */
int emu_wbfm_batch(const uint8_t *iq, uint32_t n_captures, uint64_t len_each_bytes, uint32_t tiles_per_segment,
                   float *audio, float *disc)
{
    b200::fill_fm_taps(c_fm_taps);
    b200::FmPlan pl = b200::plan_wbfm_batch(len_each_bytes, n_captures, 148);
    if (tiles_per_segment) {
        pl.tiles_per_segment = tiles_per_segment;
        pl.segments = (uint32_t)b200::ceil_div(pl.n_tiles, tiles_per_segment);
    }
    FmParams p{};
    p.iq = iq;
    p.capture_stride = len_each_bytes;
    p.capture_bytes = len_each_bytes;
    p.m1 = pl.m1;
    p.m_base = 0;
    p.n_tiles = pl.n_tiles;
    p.total_chunks = pl.total_chunks;
    p.tiles_per_segment = pl.tiles_per_segment;
    p.audio = audio;
    p.audio_stride = b200::wbfm_audio_len(len_each_bytes);
    p.audio_base = 0;
    p.disc = disc;
    p.disc_stride = pl.m1;
    p.state = nullptr;
    p.n_audio_out = nullptr;
    emu::launch(dim3(pl.segments, n_captures), dim3(B200_FM_THREADS), B200_FM_SMEM_BYTES, [&] { k_wbfm(p); });
    return (int)pl.segments;
}

/* WBFM streaming step: `n_chunks` whole 120-sample chunks starting at stream chunk index
 * `chunk_base`, exact state carried in *state (FmState, zero it before the first call). */
int emu_wbfm_stream(const uint8_t *iq, uint32_t n_chunks, uint64_t chunk_base, void *state, float *audio,
                    uint32_t *n_audio, float *disc, uint32_t tiles_per_segment)
{
    b200::fill_fm_taps(c_fm_taps);
    FmParams p{};
    p.iq = iq;
    p.capture_stride = 0;
    p.capture_bytes = (uint64_t)n_chunks * 2 * B200_FM_CHUNK;
    p.m1 = (uint64_t)n_chunks * B200_FM_OPT;
    p.m_base = chunk_base * B200_FM_OPT;
    p.total_chunks = n_chunks;
    p.n_tiles = (uint32_t)b200::ceil_div(n_chunks, B200_FM_THREADS);
    p.tiles_per_segment = tiles_per_segment ? tiles_per_segment : (p.n_tiles ? p.n_tiles : 1); /* 0: one segment */
    const uint32_t segments = p.n_tiles ? (uint32_t)b200::ceil_div(p.n_tiles, p.tiles_per_segment) : 1;
    p.audio = audio;
    p.audio_stride = 0;
    p.audio_base = b200::ceil_div(p.m_base, B200_FM_D2);
    p.disc = disc;
    p.disc_stride = 0;
    FmState next{}; /* segments run "concurrently": the state is double-buffered like in the product */
    p.state = (const FmState *)state;
    p.state_out = &next;
    emu::launch(dim3(segments, 1), dim3(B200_FM_THREADS), B200_FM_SMEM_BYTES, [&] { k_wbfm(p); });
    memcpy(state, &next, sizeof next);
    *n_audio = (uint32_t)(b200::ceil_div(p.m_base + p.m1, B200_FM_D2) - p.audio_base);
    return 0;
}

/* WBFM, tensor-core engine (csrc/wbfm_tc.cuh): the epilogue threads run on the host, the u8 x s8 -> s32 product is taken
 * from the same B image and source addressing (exact, like the hardware).  dbg_acc: optional [128][112] raw accumulators
 * of tile 0 of capture 0; q_out: optional [3][80] tap slices; returns the tap exponent e. */
int emu_wbfm_tc_batch(const uint8_t *iq, uint32_t n_captures, uint64_t len_each_bytes, uint32_t tiles_per_segment,
                      float *audio, float *disc, int32_t *dbg_acc, int8_t *q_out)
{
    static std::vector<uint8_t> image(B200_TC_B_BYTES);
    int e = 0;
    b200::fill_fm_tc(c_fm_tc, image.data(), (int8_t(*)[B200_FM_T1])q_out, &e);
    b200::FmTcPlan pl = b200::plan_wbfm_tc(len_each_bytes, n_captures, 148);
    if (tiles_per_segment) {
        pl.tiles_per_segment = tiles_per_segment;
        pl.segments = (uint32_t)b200::ceil_div(pl.n_tiles, tiles_per_segment);
    }
    FmTcParams p{};
    p.iq = iq;
    p.capture_stride = len_each_bytes;
    p.capture_bytes = len_each_bytes;
    p.m1 = pl.m1;
    p.n_tiles = pl.n_tiles;
    p.total_rows = pl.total_rows;
    p.tiles_per_segment = pl.tiles_per_segment;
    p.segments = pl.segments;
    p.n_captures = n_captures;
    p.manual_from_tile = 0;
    p.audio = audio;
    p.audio_stride = b200::wbfm_audio_len(len_each_bytes);
    p.disc = disc;
    p.disc_stride = pl.m1;
    p.b_image = image.data();
    p.error = nullptr;
    p.dbg_acc = dbg_acc;
    p.dbg_flags = 0;
    /* fewer CTAs than work items: every CTA walks several */
    const uint32_t items = pl.segments * n_captures;
    emu::launch(dim3(items > 2 ? 2 : items, 1), dim3(B200_TC_GROUPS * B200_TC_EPI), B200_TC_SMEM_BYTES, [&] { k_wbfm_tc(p); });
    return e;
}

int emu_sizeof_fm_state(void) { return (int)sizeof(FmState); }
int emu_fm_chunk(void) { return B200_FM_CHUNK; }

/* the product's spectrum launch plan (csrc/plan.h): out4 = frames, frames_per_warp, units_per_capture, grid */
void emu_plan_spectrum(uint64_t len_bytes, uint32_t n_captures, uint32_t sm_count, uint32_t *out4)
{
    const b200::SpectrumPlan pl = b200::plan_spectrum(len_bytes, n_captures, sm_count);
    out4[0] = pl.frames;
    out4[1] = pl.frames_per_warp;
    out4[2] = pl.units_per_capture;
    out4[3] = pl.grid;
}

/* K2 conversion (window may be NULL) as b200sdr_convert_cf32_dev launches it; len_bytes % 16 == 0 */
int emu_convert_cf32(const uint8_t *iq, uint64_t len_bytes, const float *window, float *out)
{
    const uint64_t n_words = len_bytes / 4;
    emu::launch(dim3((unsigned)((n_words + 1023) / 1024), 1), dim3(256), 0,
                [&] { k_convert_cf32((const uint32_t *)iq, (float4 *)out, n_words, window); });
    return 0;
}

/* the device generator of the synthetic captures as b200sdr_synth_fill_dev launches it; `lut` = the sine table */
int emu_synth(uint8_t *out, uint32_t n_captures, uint64_t len_each, uint32_t kind, uint64_t first_capture, const float *lut)
{
    const uint64_t groups = len_each / 16;
    emu::launch(dim3((unsigned)((groups + 255) / 256), n_captures), dim3(256), 0,
                [&] { k_synth((uint4 *)out, groups, groups, kind, first_capture, lut); });
    return 0;
}

/* presentation kernels as b200sdr_render_spectrum_dev / b200sdr_render_waterfall_dev launch them */
int emu_render_spectrum(const float *spectra, uint32_t n_spectra, float scale, float db_min, float db_max, uint32_t *argb)
{
    float thr[B200_LCD_H];
    b200_fill_thresholds(thr, db_min, db_max);
    emu::launch(dim3(n_spectra), dim3(B200_LCD_W), 0, [&] { k_render_spectrum(spectra, scale, thr, argb); });
    return 0;
}
int emu_render_waterfall(const float *spectra, uint32_t n_rows, float scale, float db_min, float db_max, uint32_t *argb)
{
    float thr[B200_LCD_H];
    b200_fill_thresholds(thr, db_min, db_max);
    emu::launch(dim3(B200_LCD_H), dim3(B200_LCD_W), 0, [&] { k_render_waterfall(spectra, n_rows, scale, thr, argb); });
    return 0;
}

/* K0 counter check over one capture starting at any 4-byte boundary; `state` (CounterStreamState, may be NULL)
 * makes it a streaming step with absolute position pos_base, as stream_counter() in csrc/api.cu launches it */
int emu_counter_check(const uint8_t *u, uint64_t len_bytes, int expect_first, uint64_t *n_breaks, uint64_t *first_break,
                      void *state, uint64_t pos_base)
{
    unsigned long long res[2] = {0ull, ~0ull};
    CounterParams p{};
    p.in = (const uint32_t *)u;
    p.n_words = len_bytes / 4u;
    p.stride_words = p.n_words;
    p.expect_first = expect_first;
    p.n_breaks = &res[0];
    p.first_break = &res[1];
    p.stream = (CounterStreamState *)state;
    p.pos_base = pos_base;
    const uint32_t blocks = (uint32_t)b200::ceil_div(b200::ceil_div(p.n_words, 4), 1024);
    emu::launch(dim3(blocks ? blocks : 1, 1), dim3(256), 0, [&] { k_counter_check(p); });
    *n_breaks = res[0];
    *first_break = res[1];
    return (int)blocks;
}
int emu_sizeof_counter_state(void) { return (int)sizeof(CounterStreamState); }

/* AM over whole captures: k_am_front (optionally segmented) + k_am_back */
int emu_am_batch(const uint8_t *iq, uint32_t n_captures, uint64_t len_each_bytes, uint32_t tiles_per_segment,
                 float *audio, float *env_out)
{
    b200::fill_am_taps(c_am_taps);
    b200::AmPlan pl = b200::plan_am_batch(len_each_bytes, n_captures, 148);
    if (tiles_per_segment) {
        pl.tiles_per_segment = tiles_per_segment;
        pl.segments = (uint32_t)b200::ceil_div(pl.n_tiles, tiles_per_segment);
    }
    std::vector<float> env((size_t)n_captures * pl.q_count);
    AmFrontParams p{};
    p.iq = iq;
    p.capture_stride = len_each_bytes;
    p.capture_bytes = len_each_bytes;
    p.q_count = pl.q_count;
    p.n_tiles = pl.n_tiles;
    p.total_chunks = pl.total_chunks;
    p.tiles_per_segment = pl.tiles_per_segment;
    p.env = env.data();
    p.env_stride = pl.q_count;
    p.state = nullptr;
    emu::launch(dim3(pl.segments, n_captures), dim3(B200_AM_THREADS), B200_AM_SMEM_BYTES, [&] { k_am_front(p); });
    if (env_out) memcpy(env_out, env.data(), env.size() * sizeof(float));
    AmBackParams b{};
    b.env = env.data();
    b.env_stride = pl.q_count;
    b.q_count = pl.q_count;
    b.q_base = 0;
    b.audio = audio;
    b.audio_stride = pl.audio_len;
    b.audio_base = 0;
    b.state = nullptr;
    emu::launch(dim3(n_captures), dim3(B200_AMB_THREADS), 0, [&] { k_am_back(b); });
    return (int)pl.segments;
}

/* AM streaming step over n_chunks whole 200-sample chunks starting at stream chunk `chunk_base` */
int emu_am_stream(const uint8_t *iq, uint32_t n_chunks, uint64_t chunk_base, void *fstate, void *bstate, float *audio,
                  uint32_t *n_audio, uint32_t tiles_per_segment)
{
    b200::fill_am_taps(c_am_taps);
    std::vector<float> env(n_chunks + 1);
    AmFrontParams p{};
    p.iq = iq;
    p.capture_bytes = (uint64_t)n_chunks * 2 * B200_AM_CHUNK;
    p.q_count = n_chunks;
    p.q_base = chunk_base;
    p.total_chunks = n_chunks;
    p.n_tiles = (uint32_t)b200::ceil_div(n_chunks, B200_AM_THREADS);
    p.tiles_per_segment = tiles_per_segment ? tiles_per_segment : (p.n_tiles ? p.n_tiles : 1); /* 0: one segment */
    const uint32_t segments = p.n_tiles ? (uint32_t)b200::ceil_div(p.n_tiles, p.tiles_per_segment) : 1;
    p.env = env.data();
    AmFrontState next{};
    p.state = (const AmFrontState *)fstate;
    p.state_out = &next;
    emu::launch(dim3(segments, 1), dim3(B200_AM_THREADS), B200_AM_SMEM_BYTES, [&] { k_am_front(p); });
    memcpy(fstate, &next, sizeof next);
    AmBackParams b{};
    b.env = env.data();
    b.q_count = n_chunks;
    b.q_base = chunk_base;
    b.audio = audio;
    b.audio_base = (2 * chunk_base + 2) / 3;
    b.state = (AmBackState *)bstate;
    emu::launch(dim3(1), dim3(B200_AMB_THREADS), 0, [&] { k_am_back(b); });
    *n_audio = (uint32_t)((2 * (chunk_base + n_chunks) + 2) / 3 - b.audio_base);
    return 0;
}
int emu_sizeof_am_front_state(void) { return (int)sizeof(AmFrontState); }
int emu_sizeof_am_back_state(void) { return (int)sizeof(AmBackState); }

} /* extern "C" */
