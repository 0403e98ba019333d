/*
 * wbfm_tc.cuh -- kernel K4 (FM), tensor-core engine (cfg.fir_engine = B200SDR_FIR_ENGINE_TENSOR, batched captures):
 * the same chain as wbfm.cuh -- u8 I/Q -> /10 80-tap FIR -> discriminator -> 75 us de-emphasis -> /5 FIR -> 48 kHz --
 * with the first FIR computed by the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM) as an
 * EXACT integer product, straight from the raw bytes.
 *
 * Reference anchor: the planned MCU shape is arm_fir_decimate_f32 (CMSIS/core/arm_math.h:3307); definition followed:
 * oracle/golden.c gold_wbfm().  DESIGN.md 5.2b has the arithmetic, the measurements and why this engine exists next to
 * the CUDA-core kernel (the FP32 pipe caps the scatter FIR at 0.71 of the HBM roofline; this form has no FP32 FIR).
 *
 * The FIR as a banded-Toeplitz GEMM.  A capture is cut into ROWS of 160 samples (320 bytes).  Row R of the A operand is
 * the 480 raw bytes that start 160 bytes (80 samples of history) before the row: bytes [320 R - 160, 320 R + 320) --
 * unsigned 8-bit, interleaved I,Q exactly as the dongle delivers them, no conversion, no de-interleave.  The row yields
 * the 16 stage-1 outputs y1[16 R + i], i = 0..15.  The B operand is constant: column n = (s, i, c) holds slice s of the
 * taps of output i on the bytes of component c (I or Q) and zero on the other component's bytes,
 *     B[(s, i, c)][2 kap + c'] = (c' == c) ? q_s[80 + 10 i - kap] : 0        (tap index 0..79, else 0)
 * where the float64-designed taps are cut into three signed 8-bit slices, h[t] 2^e = q0 2^-7 + q1 2^-14 + q2 2^-21
 * (21 bits + sign: 3.6e-7 of the largest tap).  D = A B is u8 x s8 -> s32, exact; one tile = 125 rows (M = 128 with three
 * idle rows: 125 x 16 = 2000 outputs = 400 audio samples, so every tile starts on an audio sample), N = 96, K = 480 =
 * 15 MMAs of K = 32 per tile.  The epilogue thread that owns TMEM lane r holds all 16 outputs of row r: it removes the
 * 127.5 offset from the leading slice exactly (an integer and a half-integer below 2^23 are exact floats), adds the two
 * small slices, and runs the 240 kS/s stages of wbfm.cuh on its 16 consecutive outputs.
 *
 * Roles (one CTA of 192 threads per SM, persistent over (capture, segment) work items):
 *   warp 0      producer: cp.async (16 bytes each, 8 lanes = 128 contiguous bytes) global -> the 128B-swizzled K-major
 *               operand layout, two tile stages; bytes before the capture / behind its end are zero-filled;
 *   warp 1      one lane issues the 15 tcgen05.mma per tile into one of two TMEM accumulator stages and commits to the
 *               "stage free" and "accumulator full" mbarriers;
 *   warps 2..5  epilogue: tcgen05.ld (warp w owns TMEM lanes 32 (w % 4) ..), slice combine, discriminator, de-emphasis
 *               scan, /5 FIR, audio stores.  The MMAs of tile t+1 run while tile t is in the epilogue.
 * Every wait is bounded (clock64): a protocol error sets *error and ends the kernel, it cannot hang the GPU.
 *
 * Host emulation (tests/emu): only the epilogue threads run; the integer product is computed from the same B image and
 * the same source addressing, so the image layout, the quantisation and everything after TMEM are checked on the CPU.
 */
#ifndef B200_WBFM_TC_CUH
#define B200_WBFM_TC_CUH

#include "wbfm.cuh"

#define B200_TC_ROW_SAMPLES 160
#define B200_TC_ROW_BYTES 320
#define B200_TC_HIST_BYTES 160                       /* 80 samples in front of the row                         */
#define B200_TC_K_BYTES 480                          /* K of the product = bytes of one A row                  */
#define B200_TC_KSTEPS 15                            /* MMAs of K = 32 per tile                                */
#define B200_TC_ROWS 125                             /* data rows per tile (TMEM lanes 125..127 idle)          */
#define B200_TC_OPR 16                               /* stage-1 outputs per row                                */
#define B200_TC_TILE_OUT (B200_TC_ROWS * B200_TC_OPR) /* 2000                                                  */
#define B200_TC_TILE_BYTES (B200_TC_ROWS * B200_TC_ROW_BYTES) /* 40000                                         */
#define B200_TC_N 96                                 /* 3 slices x 16 outputs x (I, Q)                         */
#define B200_TC_A_BOX (128 * 128)                    /* 128 rows x 128 bytes of K, 128B swizzle                */
#define B200_TC_A_STAGE (4 * B200_TC_A_BOX)
#define B200_TC_B_BOX (B200_TC_N * 128)
#define B200_TC_B_BYTES (4 * B200_TC_B_BOX)          /* 49152                                                  */
#define B200_TC_THREADS 192
#define B200_TC_EPI 128
#define B200_TC_ACC_COLS 128                         /* TMEM columns per accumulator stage (96 used)           */
#define B200_TC_APT 4

#define B200_TC_SM_A 0
#define B200_TC_SM_B (B200_TC_SM_A + 2 * B200_TC_A_STAGE)
#define B200_TC_SM_E (B200_TC_SM_B + B200_TC_B_BYTES)                              /* float [52 + 2048 + 16] */
#define B200_TC_SM_YLAST (B200_TC_SM_E + (B200_FM_HPAD + 128 * B200_TC_OPR + 16) * 4) /* c2 [132]            */
#define B200_TC_SM_WSUM (B200_TC_SM_YLAST + 132 * 8)                               /* float [8]              */
#define B200_TC_SM_YLASTC (B200_TC_SM_WSUM + 32)                                   /* c2 [2]                 */
#define B200_TC_SM_BAR (B200_TC_SM_YLASTC + 16)                                    /* u64 [8]                */
#define B200_TC_SM_MISC (B200_TC_SM_BAR + 64)                                      /* u32 [4]                */
#define B200_TC_SMEM_BYTES (B200_TC_SM_MISC + 16)

/* byte offset of element (row n, K byte k) inside an operand of `rows` rows: K-major, 128-byte swizzle
 * (cute Swizzle<3,4,3>): boxes of 128 K-bytes, row pitch 128, the 16-byte chunk index XORed with row % 8 */
#define B200_TC_OP_OFF(rows, n, k) \
    ((uint32_t)(((k) >> 7) * ((rows) * 128) + (n) * 128 + (((((k) & 127) >> 4) ^ ((n) & 7)) << 4) + ((k) & 15)))

struct FmTcConsts {
    float b0;            /* 127.5 sum q0: the offset's share of the leading slice (exact in fp32)              */
    float c0, c1, c2;    /* 2^-7 / 2^e, 2^-14 / 2^e, 2^-21 / 2^e                                               */
    float k12;           /* -127.5 (c1 sum q1 + c2 sum q2)                                                     */
    float b0_first[16];  /* the same for the first row of a capture, whose history bytes are zero-filled:      */
    float k12_first[16]; /* output i sees real samples through taps t <= 10 i only (x[n < 0] = 0)              */
    float apow[16];      /* a^(i+1)                                                                            */
    float a16pow[5];     /* (a^16)^(2^s)                                                                       */
    float a512;          /* (a^16)^32: decay over one warp                                                     */
    float a16;
    float alpha;
    float h2[B200_FM_T2];
};

struct FmTcParams {
    const uint8_t *iq;       /* capture c at iq + c * capture_stride (16-byte aligned)                          */
    uint64_t capture_stride; /* bytes                                                                           */
    uint64_t capture_bytes;  /* valid bytes per capture (multiple of 16)                                        */
    uint64_t m1;             /* stage-1 outputs per capture                                                     */
    uint32_t n_tiles;        /* tiles per capture                                                               */
    uint32_t total_rows;     /* rows per capture = ceil(samples / 160)                                          */
    uint32_t tiles_per_segment, segments, n_captures;
    float *audio;            /* [capture][audio_stride]                                                         */
    uint64_t audio_stride;
    float *disc;             /* optional [capture][disc_stride]                                                 */
    uint64_t disc_stride;
    const uint8_t *b_image;  /* B200_TC_B_BYTES: the B operand as it lies in shared memory                      */
    uint32_t *error;         /* device word, 0 = ok; else which bounded wait expired                            */
    int32_t *dbg_acc;        /* optional [128][96]: raw accumulators of tile 0 of capture 0 (tests)             */
    uint32_t dbg_flags;      /* timing experiments only: 1 = producer copies nothing, 2 = epilogue computes nothing */
};

#ifdef B200_EMULATED
static FmTcConsts c_fm_tc;
#else
__constant__ FmTcConsts c_fm_tc;
#endif

#if defined(__CUDACC__) && !defined(B200_EMULATED) /* ------------------------------------ device-only plumbing */

B200_DEV uint32_t b200_tc_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
B200_DEV void b200_tc_bar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b200_tc_smem(bar)), "r"(count) : "memory");
}

B200_DEV uint64_t b200_tc_desc(uint32_t saddr)
{
    /* cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30) = 1 (K-major swizzled: unused),
     * SBO >> 4 [32,46) = 1024 >> 4 (8 rows x 128 bytes), version 1 [46,48), layout SWIZZLE_128B = 2 [61,64) */
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
/* cute::UMMA::InstrDescriptor: D = S32 (2 at [4,6)), A = unsigned 8-bit (0 at [7,10)), B = signed 8-bit (1 at [10,13)),
 * both K-major, N >> 3 at [17,23), M >> 4 at [24,29) */
#define B200_TC_IDESC ((2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(B200_TC_N >> 3) << 17) | ((128u >> 4) << 24))

B200_DEV void b200_tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(B200_TC_IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
B200_DEV void b200_tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b200_tc_smem(bar)) : "memory");
}
B200_DEV void b200_tc_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b200_tc_smem(bar)) : "memory");
}
/* bounded wait; gives up when another role has already failed */
B200_DEV bool b200_tc_wait(uint64_t *bar, uint32_t parity, volatile uint32_t *abort_flag)
{
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(b200_tc_smem(bar)), "r"(parity)
                     : "memory");
        if (ok) return true;
        if (*abort_flag) return false;
        if (clock64() - t0 > (1ll << 28)) return false;
    }
}
B200_DEV void b200_tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
B200_DEV void b200_tc_cp16(uint32_t sdst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(src_bytes) : "memory");
}
B200_DEV void b200_tc_epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

#else

B200_DEV void b200_tc_epi_sync() { __syncthreads(); }

#endif

/* (capture, segment) -> tile range; segments > 0 pre-roll one tile with stores suppressed (discriminator / de-emphasis /
 * audio-FIR state; the stage-1 FIR has no state here: a row carries its own history bytes) */
struct FmTcItem {
    uint32_t capture, seg, t_begin, t_end, t_first_store;
};
B200_DEV FmTcItem b200_tc_item(const FmTcParams &p, uint32_t item)
{
    FmTcItem w;
    w.capture = item / p.segments;
    w.seg = item % p.segments;
    w.t_begin = w.seg * p.tiles_per_segment;
    w.t_end = w.t_begin + p.tiles_per_segment;
    if (w.t_end > p.n_tiles) w.t_end = p.n_tiles;
    w.t_first_store = w.t_begin;
    if (w.seg > 0) w.t_begin -= 1;
    return w;
}

/* The 240 kS/s stages of one tile, run by the 128 epilogue threads.  `acc` = this thread's accumulator row:
 * acc[32 s + 2 i + c].  `row` = TMEM lane = row of the tile, `et` = index among the epilogue threads, `lw` = row / 32. */
B200_DEV void b200_tc_epilogue(const FmTcParams &p, const FmTcItem &w, uint32_t tile, uint32_t it, const uint32_t (&acc)[96],
                               int row, int et, unsigned char *smem)
{
    const FmTcConsts *k = &c_fm_tc;
    float *s_e = reinterpret_cast<float *>(smem + B200_TC_SM_E);
    c2 *s_ylast = reinterpret_cast<c2 *>(smem + B200_TC_SM_YLAST);
    float *s_wsum = reinterpret_cast<float *>(smem + B200_TC_SM_WSUM);
    c2 *s_ylastc = reinterpret_cast<c2 *>(smem + B200_TC_SM_YLASTC);
    const int lane = row & 31, lw = row >> 5;
    const bool store = tile >= w.t_first_store;
    const uint64_t m0 = (uint64_t)tile * B200_TC_TILE_OUT;
    int last = (int)(p.total_rows - tile * B200_TC_ROWS) - 1; /* last row of the tile that holds samples */
    if (last > B200_TC_ROWS - 1) last = B200_TC_ROWS - 1;
    const int par = (int)(it & 1);

    /* ---- slices -> y1: (d0 - 127.5 sum q0) c0 + d1 c1 + d2 c2 - 127.5 (c1 sum q1 + c2 sum q2) ---- */
    float yr[B200_TC_OPR], yi[B200_TC_OPR];
    const bool first_row = tile == 0 && row == 0; /* history bytes are zeros, not samples */
#pragma unroll
    for (int i = 0; i < B200_TC_OPR; ++i) {
        const float b0 = first_row ? k->b0_first[i] : k->b0;
        const float k12 = first_row ? k->k12_first[i] : k->k12;
        const float r0 = (float)(int32_t)acc[2 * i] - b0, q0 = (float)(int32_t)acc[2 * i + 1] - b0;
        yr[i] = fmaf(r0, k->c0, fmaf((float)(int32_t)acc[32 + 2 * i], k->c1, fmaf((float)(int32_t)acc[64 + 2 * i], k->c2, k12)));
        yi[i] = fmaf(q0, k->c0, fmaf((float)(int32_t)acc[32 + 2 * i + 1], k->c1, fmaf((float)(int32_t)acc[64 + 2 * i + 1], k->c2, k12)));
    }
    if (row == last) s_ylastc[par] = c2_make(yr[B200_TC_OPR - 1], yi[B200_TC_OPR - 1]);
    else s_ylast[row + 1] = c2_make(yr[B200_TC_OPR - 1], yi[B200_TC_OPR - 1]);
    b200_tc_epi_sync(); /* E1 */

    /* ---- discriminator + thread-serial de-emphasis ---- */
    float e[B200_TC_OPR];
    {
        float pr, pi;
        c2_get(row == 0 ? s_ylastc[par ^ 1] : s_ylast[row], pr, pi);
        float d[B200_TC_OPR];
#pragma unroll
        for (int i = 0; i < B200_TC_OPR; ++i) {
            const float zr = fmaf(yr[i], pr, yi[i] * pi);
            const float zi = fmaf(yi[i], pr, -(yr[i] * pi));
            d[i] = b200_atan2(zi, zr);
            pr = yr[i];
            pi = yi[i];
        }
        if (m0 == 0 && row == 0) d[0] = 0.0f; /* y1[-1] = 0: defined as d[0] = 0 */
        if (p.disc && store && row < B200_TC_ROWS) { /* lanes 125..127 hold no row of this tile */
            const uint64_t m = m0 + (uint64_t)row * B200_TC_OPR;
            float *dst = p.disc + (uint64_t)w.capture * p.disc_stride + m;
            const int n_valid = p.m1 > m ? (p.m1 - m > B200_TC_OPR ? B200_TC_OPR : (int)(p.m1 - m)) : 0;
#pragma unroll
            for (int i = 0; i < B200_TC_OPR; ++i)
                if (i < n_valid) dst[i] = d[i];
        }
        const float a1 = 1.0f - k->alpha;
        float run = 0.0f;
#pragma unroll
        for (int i = 0; i < B200_TC_OPR; ++i) {
            run = fmaf(a1, run, k->alpha * d[i]);
            e[i] = run;
        }
    }
    /* warp scan of the row totals, then the carry across warps (wbfm.cuh, with 16 outputs per thread) */
    float v = e[B200_TC_OPR - 1];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const float u = __shfl_up_sync(0xffffffffu, v, 1u << s);
        if (lane >= (1 << s)) v = fmaf(k->a16pow[s], u, v);
    }
    float vprev = __shfl_up_sync(0xffffffffu, v, 1u);
    if (lane == 0) vprev = 0.0f;
    if (lane == 31) s_wsum[lw] = v;
    float lane_pow = 1.0f; /* (a^16)^lane */
#pragma unroll
    for (int s = 0; s < 5; ++s)
        if (lane & (1 << s)) lane_pow *= k->a16pow[s];
    b200_tc_epi_sync(); /* E2 */
    {
        float cw = s_wsum[4]; /* carry into warp 0 = e[m0 - 1] */
        for (int q = 0; q < lw; ++q) cw = fmaf(k->a512, cw, s_wsum[q]);
        const float cin = fmaf(lane_pow, cw, vprev);
#pragma unroll
        for (int i = 0; i < B200_TC_OPR; ++i) e[i] = fmaf(k->apow[i], cin, e[i]);
        float4 *dst = reinterpret_cast<float4 *>(s_e + B200_FM_HPAD + row * B200_TC_OPR);
#pragma unroll
        for (int q = 0; q < B200_TC_OPR / 4; ++q) dst[q] = make_float4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
    }
    b200_tc_epi_sync(); /* E3 */

    /* ---- stage 2: audio[p] = sum_k h2[k] e[5 p - k]; a tile starts at a multiple of 20 stage-1 outputs, thread et < 100
     * owns the four audio samples whose window is e[20 et - 49 .. 20 et + 15] (wbfm.cuh stage 2) ---- */
    {
        uint64_t mg_end = m0 + (uint64_t)(last + 1) * B200_TC_OPR;
        if (mg_end > p.m1) mg_end = p.m1;
        const uint64_t pg_first = m0 / B200_FM_D2;
        const uint64_t pg_end = (mg_end + B200_FM_D2 - 1) / B200_FM_D2;
        const uint64_t pg0 = pg_first + (uint64_t)et * B200_TC_APT;
        if (et < B200_TC_TILE_OUT / (B200_FM_D2 * B200_TC_APT) && pg0 < pg_end) {
            constexpr int WOFF = B200_FM_HPAD - B200_FM_HIST; /* 3 */
            constexpr int NW4 = (WOFF + B200_FM_T2 + B200_FM_D2 * (B200_TC_APT - 1) + 3) / 4; /* 17 */
            const float4 *win = reinterpret_cast<const float4 *>(s_e + et * (B200_FM_D2 * B200_TC_APT));
            float ew[4 * NW4];
#pragma unroll
            for (int j = 0; j < NW4; ++j) {
                const float4 q = win[j];
                ew[4 * j] = q.x; ew[4 * j + 1] = q.y; ew[4 * j + 2] = q.z; ew[4 * j + 3] = q.w;
            }
            float au[B200_TC_APT];
#pragma unroll
            for (int r = 0; r < B200_TC_APT; ++r) au[r] = 0.0f;
#pragma unroll
            for (int t = 0; t < B200_FM_T2; ++t) {
                const float hk = k->h2[t];
#pragma unroll
                for (int r = 0; r < B200_TC_APT; ++r) au[r] = fmaf(hk, ew[WOFF + B200_FM_HIST + B200_FM_D2 * r - t], au[r]);
            }
            if (store) {
                float *dst = p.audio + (uint64_t)w.capture * p.audio_stride + pg0;
#pragma unroll
                for (int r = 0; r < B200_TC_APT; ++r)
                    if (pg0 + r < pg_end) dst[r] = au[r];
            }
        }
    }
    b200_tc_epi_sync(); /* E4 */
    /* carries for the next tile */
    if (et == 0) s_wsum[4] = s_e[B200_FM_HPAD + (last + 1) * B200_TC_OPR - 1];
    const float hv = et < B200_FM_HIST ? s_e[B200_FM_HPAD + (last + 1) * B200_TC_OPR - B200_FM_HIST + et] : 0.0f;
    b200_tc_epi_sync(); /* E5: history source read before it is overwritten (partial tiles overlap) */
    if (et < B200_FM_HIST) s_e[B200_FM_HPAD - B200_FM_HIST + et] = hv;
    /* the next tile's E1..E2 order these writes before their readers */
}

/* zero state at the start of a work item (epilogue threads) */
B200_DEV void b200_tc_reset_state(int et, unsigned char *smem)
{
    float *s_e = reinterpret_cast<float *>(smem + B200_TC_SM_E);
    float *s_wsum = reinterpret_cast<float *>(smem + B200_TC_SM_WSUM);
    c2 *s_ylastc = reinterpret_cast<c2 *>(smem + B200_TC_SM_YLASTC);
    if (et < B200_FM_HIST) s_e[B200_FM_HPAD - B200_FM_HIST + et] = 0.0f;
    if (et == 64) s_wsum[4] = 0.0f;
    if (et == 65) s_ylastc[1] = c2_zero(); /* tile 0 of an item reads carry buffer 1 */
    b200_tc_epi_sync();
}

#ifdef B200_EMULATED
/* the integer product as the tensor cores compute it, from the same source addressing and the same B image */
static void b200_tc_emulated_acc(const FmTcParams &p, const FmTcItem &w, uint32_t tile, int row, uint32_t (&acc)[96])
{
    const uint8_t *cap = p.iq + (uint64_t)w.capture * p.capture_stride;
    const int64_t row_byte = ((int64_t)tile * B200_TC_ROWS + row) * B200_TC_ROW_BYTES - B200_TC_HIST_BYTES;
    for (int n = 0; n < B200_TC_N; ++n) {
        int64_t s = 0;
        for (int kk = 0; kk < B200_TC_K_BYTES; ++kk) {
            const int64_t off = row_byte + kk, chunk = off & ~(int64_t)15;
            const int a = (chunk >= 0 && chunk + 16 <= (int64_t)p.capture_bytes) ? cap[off] : 0;
            s += (int64_t)a * (int64_t)(int8_t)p.b_image[B200_TC_OP_OFF(B200_TC_N, n, kk)];
        }
        acc[n] = (uint32_t)(int32_t)s;
    }
}
#endif

__global__ void __launch_bounds__(B200_TC_THREADS, 1) k_wbfm_tc(FmTcParams p)
{
    B200_DYN_SMEM(smem);
    const int tid = (int)threadIdx.x;
    const uint32_t n_items = p.segments * p.n_captures;
#ifdef B200_EMULATED
    /* 128 threads: the epilogue role only */
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const FmTcItem w = b200_tc_item(p, item);
        b200_tc_reset_state(tid, smem);
        for (uint32_t tile = w.t_begin, it = 0; tile < w.t_end; ++tile, ++it) {
            uint32_t acc[96];
            if (tid < B200_TC_ROWS) b200_tc_emulated_acc(p, w, tile, tid, acc);
            else for (int n = 0; n < 96; ++n) acc[n] = 0x12345u * (uint32_t)(n + tid); /* idle lanes hold anything */
            if (p.dbg_acc && w.capture == 0 && tile == 0)
                for (int n = 0; n < 96; ++n) p.dbg_acc[tid * 96 + n] = (int32_t)acc[n];
            b200_tc_epilogue(p, w, tile, it, acc, tid, tid, smem);
        }
    }
#else
    const int warp = tid >> 5, lane = tid & 31;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + B200_TC_SM_BAR);
    uint64_t *bar_full = s_bar, *bar_empty = s_bar + 2, *bar_tfull = s_bar + 4, *bar_tempty = s_bar + 6;
    uint32_t *s_misc = reinterpret_cast<uint32_t *>(smem + B200_TC_SM_MISC); /* [0] TMEM base, [1] abort flag */
    volatile uint32_t *s_abort = s_misc + 1;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            b200_tc_bar_init(bar_full + i, 1);
            b200_tc_bar_init(bar_empty + i, 1);
            b200_tc_bar_init(bar_tfull + i, 1);
            b200_tc_bar_init(bar_tempty + i, 4); /* one arrival per epilogue warp */
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_misc[1] = 0u;
    }
    if (warp == 1) { /* TMEM: two accumulator stages of 128 columns */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(b200_tc_smem(s_misc)), "r"(2u * B200_TC_ACC_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    { /* the constant B operand */
        const uint4 *src = reinterpret_cast<const uint4 *>(p.b_image);
        uint4 *dst = reinterpret_cast<uint4 *>(smem + B200_TC_SM_B);
        for (int i = tid; i < B200_TC_B_BYTES / 16; i += B200_TC_THREADS) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> visible to the MMA's async proxy */
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_misc[0];
    const uint32_t a_base = b200_tc_smem(smem + B200_TC_SM_A), b_base = b200_tc_smem(smem + B200_TC_SM_B);

    if (warp == 0) {
        /* ===== producer ===== */
        uint32_t g = 0;
        bool pending = false, ok = true;
        uint32_t pending_stage = 0;
        const int c8 = lane & 7, rsub = lane >> 3;
        for (uint32_t item = blockIdx.x; item < n_items && ok; item += gridDim.x) {
            const FmTcItem w = b200_tc_item(p, item);
            const uint8_t *cap = p.iq + (uint64_t)w.capture * p.capture_stride;
            for (uint32_t tile = w.t_begin; tile < w.t_end; ++tile, ++g) {
                const uint32_t stage = g & 1u, phase = (g >> 1) & 1u;
                if (!b200_tc_wait(bar_empty + stage, phase ^ 1u, s_abort)) { ok = false; break; }
                const int64_t tile_byte = (int64_t)tile * B200_TC_TILE_BYTES - B200_TC_HIST_BYTES;
                const uint32_t a_stage = a_base + stage * B200_TC_A_STAGE;
#pragma unroll 1
                for (int b = 0; b < ((p.dbg_flags & 1u) ? 0 : 4); ++b) {
                    const int x = b * 128 + c8 * 16;
                    if (x < B200_TC_K_BYTES) {
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) {
                            const int r = rr * 4 + rsub;
                            if (r < B200_TC_ROWS) {
                                const int64_t off = tile_byte + (int64_t)r * B200_TC_ROW_BYTES + x;
                                const bool valid = off >= 0 && off + 16 <= (int64_t)p.capture_bytes;
                                b200_tc_cp16(a_stage + (uint32_t)(b * B200_TC_A_BOX + r * 128 + ((c8 ^ (r & 7)) << 4)), cap + (valid ? off : 0),
                                             valid ? 16u : 0u);
                            }
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (pending) { /* the tile before this one has landed: hand it to the MMA warp */
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) b200_tc_arrive(bar_full + pending_stage);
                }
                pending = true;
                pending_stage = stage;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (pending && ok) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) b200_tc_arrive(bar_full + pending_stage);
        }
        if (!ok && lane == 0) { *s_abort = 1u; atomicCAS(p.error, 0u, 1u); }
    } else if (warp == 1) {
        /* ===== MMA issuer (one lane) ===== */
        if (lane == 0) {
            uint32_t g = 0;
            uint32_t fail = 0;
            for (uint32_t item = blockIdx.x; item < n_items && !fail; item += gridDim.x) {
                const FmTcItem w = b200_tc_item(p, item);
                for (uint32_t tile = w.t_begin; tile < w.t_end; ++tile, ++g) {
                    const uint32_t stage = g & 1u, phase = (g >> 1) & 1u;
                    if (!b200_tc_wait(bar_full + stage, phase, s_abort)) { fail = 2; break; }
                    if (!b200_tc_wait(bar_tempty + stage, phase ^ 1u, s_abort)) { fail = 3; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_stage = a_base + stage * B200_TC_A_STAGE;
#pragma unroll
                    for (int ks = 0; ks < B200_TC_KSTEPS; ++ks) {
                        const uint64_t da = b200_tc_desc(a_stage + (uint32_t)((ks >> 2) * B200_TC_A_BOX + (ks & 3) * 32));
                        const uint64_t db = b200_tc_desc(b_base + (uint32_t)((ks >> 2) * B200_TC_B_BOX + (ks & 3) * 32));
                        b200_tc_mma(tmem + stage * B200_TC_ACC_COLS, da, db, ks > 0 ? 1u : 0u);
                    }
                    b200_tc_commit(bar_empty + stage);  /* the operand stage is free once the MMAs have read it */
                    b200_tc_commit(bar_tfull + stage);  /* ... and the accumulator is complete                  */
                }
            }
            if (fail) { *s_abort = 1u; atomicCAS(p.error, 0u, fail); }
        }
        __syncwarp();
    } else {
        /* ===== epilogue ===== */
        const int q = warp & 3;              /* TMEM lane quarter this warp may read */
        const int row = q * 32 + lane;
        const int et = (warp - 2) * 32 + lane;
        uint32_t g = 0;
        bool ok = true;
        for (uint32_t item = blockIdx.x; item < n_items && ok; item += gridDim.x) {
            const FmTcItem w = b200_tc_item(p, item);
            b200_tc_reset_state(et, smem);
            for (uint32_t tile = w.t_begin, it = 0; tile < w.t_end; ++tile, ++it, ++g) {
                const uint32_t stage = g & 1u, phase = (g >> 1) & 1u;
                if (!b200_tc_wait(bar_tfull + stage, phase, s_abort)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t acc[96];
                {
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + stage * B200_TC_ACC_COLS;
                    uint32_t v0[32], v1[32], v2[32];
                    b200_tc_ld32(taddr, v0);
                    b200_tc_ld32(taddr + 32, v1);
                    b200_tc_ld32(taddr + 64, v2);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int n = 0; n < 32; ++n) { acc[n] = v0[n]; acc[32 + n] = v1[n]; acc[64 + n] = v2[n]; }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) b200_tc_arrive(bar_tempty + stage); /* the MMA warp may overwrite this accumulator */
                if (p.dbg_acc && w.capture == 0 && tile == 0) {
#pragma unroll
                    for (int n = 0; n < 96; ++n) p.dbg_acc[row * 96 + n] = (int32_t)acc[n];
                }
                if (!(p.dbg_flags & 2u)) b200_tc_epilogue(p, w, tile, it, acc, row, et, smem);
            }
        }
        if (!ok && lane == 0) { *s_abort = 1u; atomicCAS(p.error, 0u, 4u); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2u * B200_TC_ACC_COLS) : "memory");
#endif
}

#endif
