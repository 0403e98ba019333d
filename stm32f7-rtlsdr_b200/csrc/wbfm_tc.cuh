/*
 * wbfm_tc.cuh -- kernel K4 (FM), tensor-core engine (cfg.fir_engine = B200SDR_FIR_ENGINE_TENSOR, batched captures):
 * the same chain as wbfm.cuh -- u8 I/Q -> /10 80-tap FIR -> discriminator -> 75 us de-emphasis -> /5 FIR -> 48 kHz --
 * with the first FIR computed by the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM) as an
 * EXACT integer product, straight from the raw bytes that TMA drops into shared memory.
 *
 * Reference anchor: the planned MCU shape is arm_fir_decimate_f32 (CMSIS/core/arm_math.h:3307); definition followed:
 * oracle/golden.c gold_wbfm().  DESIGN.md 5.2b has the arithmetic, the measurements and why this engine exists next to
 * the CUDA-core kernel (the FP32 pipe caps the scatter FIR at 0.71 of the HBM roofline; this form has no FP32 FIR).
 *
 * The FIR as a banded-Toeplitz GEMM.  A capture is cut into ROWS of 160 samples (320 bytes).  Row R of the A operand is
 * 512 raw bytes: 192 bytes (96 samples) of history, [320 R - 192, 320 R), then the row's own 320 bytes -- unsigned
 * 8-bit, interleaved I,Q exactly as the dongle delivers them, no conversion, no de-interleave.  The row yields the 16
 * stage-1 outputs y1[16 R + o], o = 0..15, and once more the output before them (o = -1), so the discriminator of a row
 * needs nothing from its neighbour.  The B operand is constant: column (s, h, j, c) holds slice s of the taps of output
 * o = 8 h - 1 + j (half-row h = 0, 1; j = 0..8) on the bytes of component c (I or Q), zero on the other component:
 *     B[(s, h, j, c)][2 kap + c'] = (c' == c) ? q_s[96 + 10 o - kap] : 0             (tap index 0..79, else 0)
 * where the float64-designed taps are cut into three signed 8-bit slices, h[t] 2^e = q0 2^-7 + q1 2^-14 + q2 2^-21
 * (21 bits + sign: 3.6e-7 of the largest tap).  D = A B is u8 x s8 -> s32, exact.  One tile = 125 rows (M = 128 with
 * three idle rows: 125 x 16 = 2000 outputs = 400 audio samples, so every tile starts on an audio sample), N = 112
 * (108 used), K = 512 = 16 MMAs of K = 32.
 *
 * Roles (one CTA of 320 threads per SM, persistent over (capture, segment) work items):
 *   warp 0      producer: ONE lane issues 8 TMA tensor copies per tile (cp.async.bulk.tensor, 64B swizzle: three
 *               64-byte columns of history from row R - 1, five of the row itself) into one of two operand stages; the
 *               tensor map is the plain (320 bytes, rows, captures) view of the batch, so the row before a capture and
 *               the rows behind its end are out of bounds = zero-filled by the hardware.  A capture whose length is
 *               not a multiple of 320 bytes has its last tile filled by the whole warp with bounds-checked cp.async
 *               instead (same layout);
 *   warp 1      one lane issues the 16 tcgen05.mma per tile into one of two TMEM accumulator stages and commits ONCE to the
 *               tile's "MMAs done" mbarrier, which both frees the operand stage and hands the accumulator on;
 *   warps 2..9  epilogue on half-rows: warp w reads TMEM lanes 32 (w % 4) .., columns of half-row (w - 2) / 4; a thread
 *               removes the 127.5 offset from the leading slice exactly (an integer and a half-integer below 2^23 are
 *               exact floats), adds the two small slices, and runs the 240 kS/s stages on its 8 consecutive outputs:
 *               discriminator, de-emphasis as a scan (thread-serial, then every warp scans the 250 half-row totals
 *               itself), /5 FIR out of a double-buffered shared window, software-pipelined: ONE CTA-wide (256-thread)
 *               barrier per tile, the scan of tile t runs next to the audio FIR of tile t-1.  The MMAs of tile t+1 run
 *               while tile t is in the epilogue.
 * Every wait is bounded (clock64): a protocol error sets *error and ends the kernel, it cannot hang the GPU.
 *
 * Host emulation (tests/emu): only the epilogue threads run; the integer product is computed from the same B image and
 * the same source addressing, so the image layout, the quantisation and everything after TMEM are checked on the CPU.
 */
#ifndef B200_WBFM_TC_CUH
#define B200_WBFM_TC_CUH

#include "wbfm.cuh"
#if defined(__CUDACC__) && !defined(B200_EMULATED)
#include <cuda.h> /* CUtensorMap */
#endif

#define B200_TC_ROW_SAMPLES 160
#define B200_TC_ROW_BYTES 320
#define B200_TC_HIST_SAMPLES 96
#define B200_TC_HIST_BYTES 192                       /* 96 samples in front of the row: 3 columns of 64 bytes   */
#define B200_TC_K_BYTES 512                          /* K of the product = bytes of one A row                   */
#define B200_TC_KSTEPS 16                            /* MMAs of K = 32 per tile                                 */
#define B200_TC_ROWS 125                             /* data rows per tile (TMEM lanes 125..127 idle)           */
#define B200_TC_OPR 16                               /* stage-1 outputs per row                                 */
#define B200_TC_OPT 8                                /* ... per epilogue thread (half a row)                    */
#define B200_TC_TILE_OUT (B200_TC_ROWS * B200_TC_OPR) /* 2000                                                   */
#define B200_TC_TILE_BYTES (B200_TC_ROWS * B200_TC_ROW_BYTES) /* 40000                                          */
#define B200_TC_N 112                                /* 3 slices x 2 half-rows x 9 outputs x (I, Q) = 108, padded */
#define B200_TC_SLICE_COLS 36
#define B200_TC_HALF_COLS 18
#define B200_TC_BOXES 8                              /* 64-byte K columns: 0..2 history, 3..7 the row           */
#define B200_TC_A_BOX (128 * 64)                     /* 128 rows x 64 bytes of K, 64B swizzle                   */
#define B200_TC_A_STAGE (B200_TC_BOXES * B200_TC_A_BOX)
#define B200_TC_B_BOX (B200_TC_N * 64)
#define B200_TC_B_BYTES (B200_TC_BOXES * B200_TC_B_BOX) /* 57344                                                */
#define B200_TC_TX_BYTES (B200_TC_BOXES * B200_TC_ROWS * 64) /* bytes one tile's TMA copies deliver             */
#define B200_TC_EPI 256
#define B200_TC_THREADS (64 + B200_TC_EPI)
#define B200_TC_ACC_COLS 128                         /* TMEM columns per accumulator stage                      */
#define B200_TC_APT 2                                 /* audio samples per thread in stage 2 (200 threads)       */
#define B200_TC_EBUF (B200_FM_HPAD + 128 * B200_TC_OPR + 16) /* floats per e[] buffer                           */

#define B200_TC_SM_A 0
#define B200_TC_SM_B (B200_TC_SM_A + 2 * B200_TC_A_STAGE)
#define B200_TC_SM_E (B200_TC_SM_B + B200_TC_B_BYTES)          /* float [2][52 + 2048 + 16]              */
#define B200_TC_SM_TOT (B200_TC_SM_E + 2 * B200_TC_EBUF * 4)   /* float [2][256]                         */
#define B200_TC_SM_CIN (B200_TC_SM_TOT + 2 * 256 * 4)          /* float [8 warps][256 + 8]               */
#define B200_TC_SM_BAR (B200_TC_SM_CIN + 8 * 264 * 4)          /* u64 [8]                                */
#define B200_TC_SM_MISC (B200_TC_SM_BAR + 64)                  /* u32 [4]                                */
#define B200_TC_SMEM_BYTES (B200_TC_SM_MISC + 16)

/* byte offset of element (row n, K byte k) inside an operand of `rows` rows: K-major, 64-byte swizzle
 * (cute Swizzle<2,4,3>): columns of 64 K-bytes, row pitch 64, the 16-byte chunk index XORed with (row / 2) % 4 */
#define B200_TC_OP_OFF(rows, n, k) \
    ((uint32_t)(((k) >> 6) * ((rows) * 64) + (n) * 64 + (((((k) & 63) >> 4) ^ (((n) >> 1) & 3)) << 4) + ((k) & 15)))
/* accumulator column of (slice s, half-row h, output j of the half-row's nine, component c) */
#define B200_TC_COL(s, h, j, c) (B200_TC_SLICE_COLS * (s) + B200_TC_HALF_COLS * (h) + 2 * (j) + (c))

struct FmTcConsts {
    float b0;            /* 127.5 sum q0: the offset's share of the leading slice (exact in fp32)              */
    float c0, c1, c2;    /* 2^-7 / 2^e, 2^-14 / 2^e, 2^-21 / 2^e                                               */
    float k12;           /* -127.5 (c1 sum q1 + c2 sum q2)                                                     */
    float b0_first[16];  /* the same for the first row of a capture, whose history bytes are zero-filled:      */
    float k12_first[16]; /* output o sees real samples through taps t <= 10 o only (x[n < 0] = 0)              */
    float apow[8];       /* a^(i+1)                                                                            */
    float a8p[9];        /* (a^8)^j, j = 0..8                                                                  */
    float a64pow[5];     /* (a^64)^(2^s): one lane of the totals scan covers 8 half-rows = 64 outputs          */
    float a8;
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
    uint32_t manual_from_tile; /* tiles >= this are filled with cp.async instead of TMA (0: no tensor map at all) */
    float *audio;            /* [capture][audio_stride]                                                         */
    uint64_t audio_stride;
    float *disc;             /* optional [capture][disc_stride]                                                 */
    uint64_t disc_stride;
    const uint8_t *b_image;  /* B200_TC_B_BYTES: the B operand as it lies in shared memory                      */
    uint32_t *error;         /* device word, 0 = ok; else which bounded wait expired                            */
    int32_t *dbg_acc;        /* optional [128][112]: raw accumulators of tile 0 of capture 0 (tests)            */
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
     * SBO >> 4 [32,46) = 512 >> 4 (8 rows x 64 bytes), version 1 [46,48), layout SWIZZLE_64B = 4 [61,64) */
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
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
B200_DEV void b200_tc_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b200_tc_smem(bar)), "r"(bytes) : "memory");
}
/* one box of the (320 bytes, rows, captures) view: 64 bytes x 125 rows, 64B swizzle, zero fill out of bounds */
B200_DEV void b200_tc_tma_box(uint32_t sdst, const void *tmap, int32_t x, int32_t row, int32_t capture, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(sdst),
                 "l"(tmap), "r"(x), "r"(row), "r"(capture), "r"(b200_tc_smem(bar))
                 : "memory");
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
B200_DEV void b200_tc_epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

#else

B200_DEV void b200_tc_epi_sync() { __syncthreads(); }

#endif

/* (capture, segment) -> tile range; segments > 0 pre-roll one tile with stores suppressed (de-emphasis / audio-FIR
 * state; the stage-1 FIR and the discriminator have no state here: a row carries its own history bytes) */
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

/* source byte offset (from the capture start) of the 16-byte chunk `chunk` of K column `box` of capture row R, and whether
 * it lies inside the capture: columns 0..2 are the last 192 bytes of row R - 1, columns 3..7 the 320 bytes of row R.
 * This is exactly what the tensor map's bounds give the TMA copies (whole rows) -- refined to 16 bytes for ragged ends. */
B200_DEV bool b200_tc_src(int64_t R, int box, int chunk, uint64_t capture_bytes, int64_t &off)
{
    if (box < 3) {
        off = (R - 1) * B200_TC_ROW_BYTES + 128 + 64 * box + 16 * chunk;
        return R >= 1 && off + 16 <= (int64_t)capture_bytes;
    }
    off = R * B200_TC_ROW_BYTES + 64 * (box - 3) + 16 * chunk;
    return off + 16 <= (int64_t)capture_bytes;
}

/* ---- The 240 kS/s stages, run by the 256 epilogue threads in three phases that are software-pipelined over the tiles
 * of a work item with ONE CTA-wide barrier per tile:
 *     A(t)  accumulators -> y1 -> discriminator -> thread-serial de-emphasis of the thread's 8 outputs (registers),
 *           half-row total -> s_tot[t & 1]
 *     ---- barrier ----
 *     C(t-1) /5 FIR of the PREVIOUS tile out of s_e[(t-1) & 1] (complete since the barrier)
 *     B1(t) every warp scans the 250 half-row totals itself -> carried-in values in its scratch   } the TMEM loads of tile
 *     A(t+1)                                                                                     } t+1 are in flight
 *     B2(t) adds the carried-in part to the thread's e[] -> s_e[t & 1]                            } under B1
 * so the long dependent chain of the scan hides behind the accumulator loads and the arithmetic of the next tile, and
 * nothing waits on a second barrier.  acc[32 s + 2 j + c] = column B200_TC_COL(s, h, j, c) as loaded.  row = TMEM lane, h = half-row,
 * et = index among the epilogue threads, ws = epilogue warp index (scratch slot). ---- */
struct FmTcTile {
    uint32_t tile, it;
    int last;      /* last row of the tile that holds samples */
    bool store;
    uint64_t m0;
};
B200_DEV FmTcTile b200_tc_tile(const FmTcParams &p, const FmTcItem &w, uint32_t tile, uint32_t it)
{
    FmTcTile t;
    t.tile = tile;
    t.it = it;
    t.store = tile >= w.t_first_store;
    t.m0 = (uint64_t)tile * B200_TC_TILE_OUT;
    t.last = (int)(p.total_rows - tile * B200_TC_ROWS) - 1;
    if (t.last > B200_TC_ROWS - 1) t.last = B200_TC_ROWS - 1;
    return t;
}

/* atan2 as in wbfm.cuh with a degree-13 odd polynomial (7 coefficients, 3.2e-7 rad in fp32) */
B200_DEV float b200_tc_atan2(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f), mn = fminf(ax, ay);
    const float a = mn * b200_rcp_fast(mx);
    const float s = a * a;
    float p = 6.811790634e-03f;
    p = fmaf(p, s, -3.360421286e-02f);
    p = fmaf(p, s, 7.962366400e-02f);
    p = fmaf(p, s, -1.323334168e-01f);
    p = fmaf(p, s, 1.980781547e-01f);
    p = fmaf(p, s, -3.331736805e-01f);
    p = fmaf(p, s, 9.999961115e-01f);
    float r = p * a;
    if (ay > ax) r = 1.57079632679489662f - r;
    if (x < 0.0f) r = 3.14159265358979324f - r;
    return copysignf(r, y);
}

/* FIRST: the tile holds the first row of a capture (tile 0), whose history bytes are zeros, not samples */
template <bool FIRST>
B200_DEV void b200_tc_phase_a(const FmTcParams &p, const FmTcItem &w, const FmTcTile &t, const uint32_t (&acc)[96], int row, int h,
                              float (&e)[B200_TC_OPT], unsigned char *smem)
{
    const FmTcConsts *k = &c_fm_tc;
    float *s_tot = reinterpret_cast<float *>(smem + B200_TC_SM_TOT) + (t.it & 1u) * 256;
    const int idx = 2 * row + h; /* position of this half-row in the tile */
    /* slices -> y1 = (d0 - 127.5 sum q0) c0 + (128 d1 + d2) c2 - 127.5 (c1 sum q1 + c2 sum q2): the leading slice loses
     * its offset exactly (integer minus half-integer, both below 2^23), the two small slices are joined as integers */
    float yr[B200_TC_OPT + 1], yi[B200_TC_OPT + 1]; /* [0] = the output before this half-row */
    const bool first_row = FIRST && row == 0;
    const float c0 = k->c0, c2 = k->c2;
#pragma unroll
    for (int j = 0; j <= B200_TC_OPT; ++j) {
        const int o = 8 * h - 1 + j;
        float b0 = k->b0, k12 = k->k12;
        if (FIRST && first_row && o >= 0) { b0 = k->b0_first[o < 0 ? 0 : o]; k12 = k->k12_first[o < 0 ? 0 : o]; }
        const float r0 = (float)(int32_t)acc[2 * j] - b0, q0 = (float)(int32_t)acc[2 * j + 1] - b0;
        const float r12 = (float)((int32_t)acc[32 + 2 * j] * 128 + (int32_t)acc[64 + 2 * j]);
        const float q12 = (float)((int32_t)acc[32 + 2 * j + 1] * 128 + (int32_t)acc[64 + 2 * j + 1]);
        yr[j] = fmaf(r0, c0, fmaf(r12, c2, k12));
        yi[j] = fmaf(q0, c0, fmaf(q12, c2, k12));
    }
    float d[B200_TC_OPT];
#pragma unroll
    for (int i = 0; i < B200_TC_OPT; ++i) {
        const float zr = fmaf(yr[i + 1], yr[i], yi[i + 1] * yi[i]);
        const float zi = fmaf(yi[i + 1], yr[i], -(yr[i + 1] * yi[i]));
        d[i] = b200_tc_atan2(zi, zr);
    }
    if (FIRST && idx == 0) d[0] = 0.0f; /* y1[-1] = 0: defined as d[0] = 0 */
    if (p.disc && t.store && row < B200_TC_ROWS) { /* lanes 125..127 hold no row of this tile */
        const uint64_t m = t.m0 + (uint64_t)idx * B200_TC_OPT;
        float *dst = p.disc + (uint64_t)w.capture * p.disc_stride + m;
        const int n_valid = p.m1 > m ? (p.m1 - m > B200_TC_OPT ? B200_TC_OPT : (int)(p.m1 - m)) : 0;
#pragma unroll
        for (int i = 0; i < B200_TC_OPT; ++i)
            if (i < n_valid) dst[i] = d[i];
    }
    const float alpha = k->alpha, a1 = 1.0f - alpha;
    float run = 0.0f;
#pragma unroll
    for (int i = 0; i < B200_TC_OPT; ++i) {
        run = fmaf(a1, run, alpha * d[i]);
        e[i] = run;
    }
    s_tot[idx] = run;
}

/* after the tile's barrier: S[i] = a^8 S[i-1] + tot[i], S[-1] = cw (e[] just before the tile, carried in a register by
 * every thread); cin[i] = S[i-1] = e[] just before half-row i; lane l of every warp takes half-rows 8 l .. 8 l + 7 */
B200_DEV void b200_tc_phase_b1(const FmTcTile &t, int et, int ws, float cw, unsigned char *smem)
{
    const FmTcConsts *k = &c_fm_tc;
    float *s_e_cur = reinterpret_cast<float *>(smem + B200_TC_SM_E) + (t.it & 1u) * B200_TC_EBUF;
    const float *s_e_prev = reinterpret_cast<const float *>(smem + B200_TC_SM_E) + ((t.it & 1u) ^ 1u) * B200_TC_EBUF;
    const float *s_tot = reinterpret_cast<const float *>(smem + B200_TC_SM_TOT) + (t.it & 1u) * 256;
    float *s_cin = reinterpret_cast<float *>(smem + B200_TC_SM_CIN) + ws * 264;
    const int lane = et & 31;
    /* the 49 (52) newest e[] of the previous tile go in front of this tile's buffer (zeros at the start of a work item;
     * the previous tile of an item is always a full one and complete since the barrier) */
    if (et < B200_FM_HPAD) s_e_cur[et] = t.it ? s_e_prev[B200_TC_TILE_OUT + et] : 0.0f;
    const float4 ta = reinterpret_cast<const float4 *>(s_tot)[2 * lane], tb = reinterpret_cast<const float4 *>(s_tot)[2 * lane + 1];
    const float tt[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
    float P[8], run = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        run = fmaf(k->a8, run, tt[j]);
        P[j] = run;
    }
    float v = run;
#pragma unroll
    for (int s = 0; s < 3; ++s) { /* 8 lanes = 512 outputs back: what lies further has decayed by a^512 = 4e-13 */
        const float u = __shfl_up_sync(0xffffffffu, v, 1u << s);
        if (lane >= (1 << s)) v = fmaf(k->a64pow[s], u, v);
    }
    float vprev = __shfl_up_sync(0xffffffffu, v, 1u);
    if (lane == 0) vprev = 0.0f;
    float lane_pow = 1.0f; /* (a^64)^lane */
#pragma unroll
    for (int s = 0; s < 5; ++s)
        if (lane & (1 << s)) lane_pow *= k->a64pow[s];
    const float lin = fmaf(lane_pow, cw, vprev); /* S[8 lane - 1] */
    float c[8];
    c[0] = lin;
#pragma unroll
    for (int j = 1; j < 8; ++j) c[j] = fmaf(k->a8p[j], lin, P[j - 1]);
    float4 *dst = reinterpret_cast<float4 *>(s_cin + 8 * lane);
    dst[0] = make_float4(c[0], c[1], c[2], c[3]);
    dst[1] = make_float4(c[4], c[5], c[6], c[7]);
}
/* ... second part: add the carried-in part to the thread's e[] -> s_e[t & 1] */
B200_DEV void b200_tc_phase_b2(const FmTcTile &t, float (&e)[B200_TC_OPT], int row, int h, int ws, float &cw, unsigned char *smem)
{
    const FmTcConsts *k = &c_fm_tc;
    float *s_e_cur = reinterpret_cast<float *>(smem + B200_TC_SM_E) + (t.it & 1u) * B200_TC_EBUF;
    const float *s_cin = reinterpret_cast<const float *>(smem + B200_TC_SM_CIN) + ws * 264;
    const int idx = 2 * row + h;
    __syncwarp();
    const float cin = s_cin[idx];
    cw = s_cin[2 * t.last + 2]; /* S[2 last + 1] = e[] after the last row that holds samples (2 last + 2 <= 250) */
    __syncwarp();               /* the scratch is rewritten by this warp in the next tile */
#pragma unroll
    for (int i = 0; i < B200_TC_OPT; ++i) e[i] = fmaf(k->apow[i], cin, e[i]);
    float4 *de = reinterpret_cast<float4 *>(s_e_cur + B200_FM_HPAD + idx * B200_TC_OPT);
    de[0] = make_float4(e[0], e[1], e[2], e[3]);
    de[1] = make_float4(e[4], e[5], e[6], e[7]);
}

/* stage 2 of a tile whose e[] is complete: audio[p] = sum_k h2[k] e[5 p - k].  A tile starts at a multiple of 10 stage-1
 * outputs; thread et < 200 owns the two audio samples p = m0 / 5 + 2 et + r whose window e[10 et - 49 .. 10 et + 5] is read
 * with 28 aligned 64-bit loads from one float in front of it. */
B200_DEV void b200_tc_phase_c(const FmTcParams &p, const FmTcItem &w, const FmTcTile &t, int et, unsigned char *smem)
{
    const FmTcConsts *k = &c_fm_tc;
    const float *s_e = reinterpret_cast<const float *>(smem + B200_TC_SM_E) + (t.it & 1u) * B200_TC_EBUF;
    uint64_t mg_end = t.m0 + (uint64_t)(t.last + 1) * B200_TC_OPR;
    if (mg_end > p.m1) mg_end = p.m1;
    const uint64_t pg_first = t.m0 / B200_FM_D2;
    const uint64_t pg_end = (mg_end + B200_FM_D2 - 1) / B200_FM_D2;
    const uint64_t pg0 = pg_first + (uint64_t)et * B200_TC_APT;
    if (et < B200_TC_TILE_OUT / (B200_FM_D2 * B200_TC_APT) && pg0 < pg_end) {
        constexpr int WOFF = 1;                                                              /* ew[j] = e[10 et - 50 + j] */
        constexpr int NW2 = (WOFF + B200_FM_T2 + B200_FM_D2 * (B200_TC_APT - 1) + 1) / 2;    /* 28 */
        const float2 *win = reinterpret_cast<const float2 *>(s_e + B200_FM_HPAD - B200_FM_HIST - WOFF + et * (B200_FM_D2 * B200_TC_APT));
        float ew[2 * NW2];
#pragma unroll
        for (int j = 0; j < NW2; ++j) {
            const float2 q = win[j];
            ew[2 * j] = q.x; ew[2 * j + 1] = q.y;
        }
        float au[B200_TC_APT];
#pragma unroll
        for (int r = 0; r < B200_TC_APT; ++r) au[r] = 0.0f;
#pragma unroll
        for (int tp = 0; tp < B200_FM_T2; ++tp) {
            const float hk = k->h2[tp];
#pragma unroll
            for (int r = 0; r < B200_TC_APT; ++r) au[r] = fmaf(hk, ew[WOFF + B200_FM_HIST + B200_FM_D2 * r - tp], au[r]);
        }
        if (t.store) {
            float *dst = p.audio + (uint64_t)w.capture * p.audio_stride + pg0;
#pragma unroll
            for (int r = 0; r < B200_TC_APT; ++r)
                if (pg0 + r < pg_end) dst[r] = au[r];
        }
    }
}

#ifdef B200_EMULATED
/* the integer product as the tensor cores compute it, from the same source addressing and the same B image */
static void b200_tc_emulated_acc(const FmTcParams &p, const FmTcItem &w, uint32_t tile, int row, int h, uint32_t (&acc)[96])
{
    const uint8_t *cap = p.iq + (uint64_t)w.capture * p.capture_stride;
    uint8_t a[B200_TC_K_BYTES];
    for (int kk = 0; kk < B200_TC_K_BYTES; ++kk) {
        int64_t off;
        const bool valid = b200_tc_src((int64_t)tile * B200_TC_ROWS + row, kk >> 6, (kk & 63) >> 4, p.capture_bytes, off);
        a[kk] = valid ? cap[off + (kk & 15)] : 0;
    }
    for (int n = 0; n < 96; ++n) acc[n] = 0xdeadu;
    for (int s = 0; s < 3; ++s)
        for (int jc = 0; jc < 32; ++jc) { /* the x32 load: 32 consecutive columns from B200_TC_COL(s, h, 0, 0) */
            const int n = B200_TC_SLICE_COLS * s + B200_TC_HALF_COLS * h + jc;
            if (n >= B200_TC_N) continue;
            int64_t sum = 0;
            for (int kk = 0; kk < B200_TC_K_BYTES; ++kk) sum += (int64_t)a[kk] * (int64_t)(int8_t)p.b_image[B200_TC_OP_OFF(B200_TC_N, n, kk)];
            acc[32 * s + jc] = (uint32_t)(int32_t)sum;
        }
}
#endif

#if defined(__CUDACC__) && !defined(B200_EMULATED)
__global__ void __launch_bounds__(B200_TC_THREADS, 1) k_wbfm_tc(FmTcParams p, const __grid_constant__ CUtensorMap tmap)
#else
__global__ void __launch_bounds__(B200_TC_THREADS, 1) k_wbfm_tc(FmTcParams p)
#endif
{
    B200_DYN_SMEM(smem);
    const int tid = (int)threadIdx.x;
    const uint32_t n_items = p.segments * p.n_captures;
#ifdef B200_EMULATED
    /* 256 threads: the epilogue role only */
    const int row = tid & 127, h = tid >> 7;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const FmTcItem w = b200_tc_item(p, item);
        float cw = 0.0f;
        float e_cur[B200_TC_OPT], e_next[B200_TC_OPT];
        auto phase_a = [&](uint32_t tile, uint32_t it, float (&e)[B200_TC_OPT]) {
            const FmTcTile t = b200_tc_tile(p, w, tile, it);
            uint32_t acc[96];
            if (row < B200_TC_ROWS) b200_tc_emulated_acc(p, w, tile, row, h, acc);
            else for (int n = 0; n < 96; ++n) acc[n] = 0x12345u * (uint32_t)(n + tid); /* idle lanes hold anything */
            if (p.dbg_acc && w.capture == 0 && tile == 0)
                for (int s = 0; s < 3; ++s)
                    for (int jc = 0; jc < B200_TC_HALF_COLS; ++jc)
                        p.dbg_acc[row * B200_TC_N + B200_TC_SLICE_COLS * s + B200_TC_HALF_COLS * h + jc] = (int32_t)acc[32 * s + jc];
            if (tile == 0) b200_tc_phase_a<true>(p, w, t, acc, row, h, e, smem);
            else b200_tc_phase_a<false>(p, w, t, acc, row, h, e, smem);
        };
        if (w.t_end > w.t_begin) phase_a(w.t_begin, 0, e_cur);
        for (uint32_t tile = w.t_begin, it = 0; tile < w.t_end; ++tile, ++it) {
            const FmTcTile t = b200_tc_tile(p, w, tile, it);
            b200_tc_epi_sync();
            if (it) b200_tc_phase_c(p, w, b200_tc_tile(p, w, tile - 1, it - 1), tid, smem);
            b200_tc_phase_b1(t, tid, tid >> 5, cw, smem);
            if (tile + 1 < w.t_end) phase_a(tile + 1, it + 1, e_next);
            b200_tc_phase_b2(t, e_cur, row, h, tid >> 5, cw, smem);
            for (int i = 0; i < B200_TC_OPT; ++i) e_cur[i] = e_next[i];
        }
        if (w.t_end > w.t_begin) {
            b200_tc_epi_sync();
            b200_tc_phase_c(p, w, b200_tc_tile(p, w, w.t_end - 1, w.t_end - 1 - w.t_begin), tid, smem);
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
            b200_tc_bar_init(bar_tempty + i, B200_TC_EPI / 32); /* one arrival per epilogue warp */
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_misc[1] = 0u;
        if (p.manual_from_tile) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
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
        bool ok = true;
        const int chunk = lane & 3, rsub = lane >> 2; /* manual fill: 8 rows x 4 chunks of one 64-byte column per instruction */
        for (uint32_t item = blockIdx.x; item < n_items && ok; item += gridDim.x) {
            const FmTcItem w = b200_tc_item(p, item);
            const uint8_t *cap = p.iq + (uint64_t)w.capture * p.capture_stride;
            for (uint32_t tile = w.t_begin; tile < w.t_end; ++tile, ++g) {
                const uint32_t stage = g & 1u, phase = (g >> 1) & 1u;
                if (!b200_tc_wait(bar_tfull + stage, phase ^ 1u, s_abort)) { ok = false; break; } /* MMAs of the tile two back done */
                const uint32_t a_stage = a_base + stage * B200_TC_A_STAGE;
                const int32_t R0 = (int32_t)(tile * B200_TC_ROWS);
                if (p.dbg_flags & 1u) {
                    if (lane == 0) b200_tc_arrive(bar_full + stage);
                } else if (tile < p.manual_from_tile) {
                    if (lane == 0) {
                        b200_tc_arrive_expect_tx(bar_full + stage, B200_TC_TX_BYTES);
#pragma unroll
                        for (int b = 0; b < B200_TC_BOXES; ++b)
                            b200_tc_tma_box(a_stage + b * B200_TC_A_BOX, &tmap, b < 3 ? 128 + 64 * b : 64 * (b - 3), b < 3 ? R0 - 1 : R0,
                                            (int32_t)w.capture, bar_full + stage);
                    }
                } else { /* ragged end of a capture (or no tensor map): bounds-checked 16-byte copies, same layout */
#pragma unroll 1
                    for (int b = 0; b < B200_TC_BOXES; ++b) {
#pragma unroll 2
                        for (int rr = 0; rr < 16; ++rr) {
                            const int r = rr * 8 + rsub;
                            if (r < B200_TC_ROWS) {
                                int64_t off;
                                const bool valid = b200_tc_src((int64_t)R0 + r, b, chunk, p.capture_bytes, off);
                                b200_tc_cp16(a_stage + (uint32_t)(b * B200_TC_A_BOX + r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4)),
                                             cap + (valid ? off : 0), valid ? 16u : 0u);
                            }
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) b200_tc_arrive(bar_full + stage);
                }
                __syncwarp();
            }
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
                    const int ksteps = (p.dbg_flags >> 8) ? (int)(p.dbg_flags >> 8) : B200_TC_KSTEPS; /* timing experiments */
#pragma unroll
                    for (int ks = 0; ks < B200_TC_KSTEPS; ++ks) {
                        if (ks >= ksteps) break;
                        const uint64_t da = b200_tc_desc(a_stage + (uint32_t)((ks >> 1) * B200_TC_A_BOX + (ks & 1) * 32));
                        const uint64_t db = b200_tc_desc(b_base + (uint32_t)((ks >> 1) * B200_TC_B_BOX + (ks & 1) * 32));
                        b200_tc_mma(tmem + stage * B200_TC_ACC_COLS, da, db, ks > 0 ? 1u : 0u);
                    }
                    /* ONE commit per tile (each costs the tensor pipe a drain, profiles/r2_wbfm_tc.txt): the same event frees
                     * the operand stage for the producer and hands the accumulator to the epilogue */
                    b200_tc_commit(bar_tfull + stage);
                }
            }
            if (fail) { *s_abort = 1u; atomicCAS(p.error, 0u, fail); }
        }
        __syncwarp();
    } else {
        /* ===== epilogue ===== */
        const int q = warp & 3;              /* TMEM lane quarter this warp may read */
        const int h = (warp - 2) >> 2;       /* half-row */
        const int row = q * 32 + lane;
        const int et = (warp - 2) * 32 + lane;
        uint32_t g = 0;
        bool ok = true;
        /* accumulator columns of this thread: issue the three TMEM loads of the tile with running index g */
        uint32_t v0[32], v1[32], v2[32];
        auto acc_issue = [&](uint32_t gi) -> bool {
            const uint32_t stage = gi & 1u, phase = (gi >> 1) & 1u;
            if (!b200_tc_wait(bar_tfull + stage, phase, s_abort)) return false;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + stage * B200_TC_ACC_COLS + B200_TC_HALF_COLS * h;
            b200_tc_ld32(taddr, v0);
            b200_tc_ld32(taddr + B200_TC_SLICE_COLS, v1);
            b200_tc_ld32(taddr + 2 * B200_TC_SLICE_COLS, v2);
            return true;
        };
        /* ... wait for them, release the accumulator stage, run phase A */
        auto acc_use = [&](const FmTcItem &w, uint32_t gi, uint32_t tile, uint32_t it, float (&e)[B200_TC_OPT]) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t acc[96];
#pragma unroll
            for (int n = 0; n < 32; ++n) { acc[n] = v0[n]; acc[32 + n] = v1[n]; acc[64 + n] = v2[n]; }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) b200_tc_arrive(bar_tempty + (gi & 1u)); /* the MMA warp may overwrite this accumulator */
            if (p.dbg_acc && w.capture == 0 && tile == 0) {
#pragma unroll
                for (int s = 0; s < 3; ++s)
#pragma unroll
                    for (int jc = 0; jc < B200_TC_HALF_COLS; ++jc)
                        p.dbg_acc[row * B200_TC_N + B200_TC_SLICE_COLS * s + B200_TC_HALF_COLS * h + jc] = (int32_t)acc[32 * s + jc];
            }
            if (p.dbg_flags & 2u) return;
            const FmTcTile t = b200_tc_tile(p, w, tile, it);
            if (tile == 0) b200_tc_phase_a<true>(p, w, t, acc, row, h, e, smem);
            else b200_tc_phase_a<false>(p, w, t, acc, row, h, e, smem);
        };
        for (uint32_t item = blockIdx.x; item < n_items && ok; item += gridDim.x) {
            const FmTcItem w = b200_tc_item(p, item);
            if (w.t_end <= w.t_begin) continue;
            float cw = 0.0f;
            float e_cur[B200_TC_OPT], e_next[B200_TC_OPT];
            if (!acc_issue(g)) { ok = false; break; }
            acc_use(w, g, w.t_begin, 0, e_cur);
            ++g;
            for (uint32_t tile = w.t_begin, it = 0; tile < w.t_end; ++tile, ++it) {
                const bool more = tile + 1 < w.t_end;
                if (p.dbg_flags & 2u) { /* timing experiment: accumulator traffic only */
                    if (more) { if (!acc_issue(g)) { ok = false; break; } acc_use(w, g, tile + 1, it + 1, e_next); ++g; }
                    continue;
                }
                const FmTcTile t = b200_tc_tile(p, w, tile, it);
                b200_tc_epi_sync();
                if (it) b200_tc_phase_c(p, w, b200_tc_tile(p, w, tile - 1, it - 1), et, smem);
                if (more && !acc_issue(g)) { ok = false; break; } /* the next tile's accumulators are on their way ... */
                b200_tc_phase_b1(t, et, warp - 2, cw, smem);        /* ... under the scan of this tile's totals        */
                if (more) { acc_use(w, g, tile + 1, it + 1, e_next); ++g; }
                b200_tc_phase_b2(t, e_cur, row, h, warp - 2, cw, smem);
#pragma unroll
                for (int i = 0; i < B200_TC_OPT; ++i) e_cur[i] = e_next[i];
            }
            if (ok && !(p.dbg_flags & 2u)) { /* drain: the audio of the item's last tile */
                b200_tc_epi_sync();
                b200_tc_phase_c(p, w, b200_tc_tile(p, w, w.t_end - 1, w.t_end - 1 - w.t_begin), et, smem);
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
