/*
 * spec_tc_spike.cuh -- EXPERIMENT, not part of the library (measured no-go, profiles/r2_spectrum_tc_spike.txt, DESIGN.md
 * 5.0b): kernel K3 with the FIRST of the two FFT-32 passes, the conversion and the window's arithmetic taken off the
 * FP32 pipe: u8 I/Q -> Hann -> 1024-point FFT -> |X|^2 -> mean / EMA, pass 1 as a tcgen05.mma kind::f16 product.
 * Correct (pass-1 accumulators within 2.6e-7 of float64, spectra within 1.4e-5 of the FP32 engine) and SLOWER than
 * csrc/spectrum.cuh: 0.65-0.70 x as one CTA of four warps, 0.44 x warp-specialised -- the per-frame chain of TMEM round
 * trips and mbarrier hand-offs cannot be hidden with two warps per scheduler, and the register file allows no more.
 *
 * Reference anchor: the planned MCU shape is arm_cfft_f32(1024) + arm_cmplx_mag_squared_f32 (CMSIS/core/arm_math.h:2149,
 * :4693; README.md:31-32); definition followed: oracle/golden.c gold_spectrum().  DESIGN.md 5.1b has the arithmetic, the
 * error analysis and the measurements.
 *
 * 1024 = 32 x 32 four-step, n = 32 a + b, k = c + 32 d:
 *     pass 1   Y0_b[c] = sum_a x[32 a + b] W32^{a c}                      (no window)            -> TENSOR CORES
 *     twiddle  Z_b[c]  = 1/4 W1024^{b c} Y0_b[c]                                                   FP32 pipe
 *     window   Yw_b[c] = 2 Z_b[c] - (Z_b[c - 1] + Z_b[c + 1])            (Hann in the DFT-32 domain, see below)
 *     pass 2   X[c + 32 d] = sum_b Yw_b[c] W32^{b d}                      register FFT-32 (fft32.cuh) after one
 *                                                                         transpose through a warp-private shared tile
 * Periodic Hann is w[n] = 1/2 - 1/4 e^{+i theta n} - 1/4 e^{-i theta n}, theta = 2 pi / 1024; multiplying by e^{-+i theta n}
 * shifts the 1024-point spectrum by one bin, and in four-step form S[k] = sum_b W1024^{b k} Y0_b[k mod 32], so the
 * window is the three-tap (1/2, -1/4, -1/4) along c of the TWIDDLED pass-1 output, with Z[-1] = W32^{-b} Z[31] and
 * Z[32] = W32^{b} Z[0] at the ends.  Two packed operations per point instead of a multiply per INPUT sample, and pass 1
 * then has a CONSTANT matrix: a contraction the tensor cores can take.
 *
 * Pass 1 on the tensor cores, exact operands.  tcgen05.mma kind::f16, M = 128, N = 64, K = 64, FP32 accumulators in TMEM:
 *     A (TMEM)   row (w, b) = lane 32 w + b holds the 32 samples x[32 a + b] of ONE frame as fp16 pairs (I - 128, Q - 128):
 *                integers of at most 8 bits, exact in fp16.  The thread that owns the TMEM lane builds its row itself:
 *                32 two-byte loads (the same 64-byte-per-warp pattern as spectrum.cuh), one PRMT (byte -> 0x6400 | u =
 *                1024 + u as fp16) and one packed half add (-1152) per complex sample, one tcgen05.st of 32 registers.
 *     B (smem)   the DFT-32 matrix as real numbers, [n = (c, re/im)][k = (a, I/Q)], as fp16 hi + fp16 lo (22 bits):
 *                two accumulating MMAs per K step.  The products are exact in fp32 (8-bit x 11-bit).
 *     D (TMEM)   lane (w, b), columns (c, re/im): thread (w, b) reads back Y0_b[c], c = 0..31, as 32 packed pairs.
 * The half-sample offset (127.5 vs 128) is a DC term 1/2 (1 + i) on every sample: it lands in column c = 0 only, as the
 * constant 16 (1 + i), and is added back there (one packed add per frame and lane, exact).
 * Eight frames make a TILE: the MMA of the even half-tile carries frames F, F+2, F+4, F+6 (one per warp), the odd one
 * F+1 .. F+7, so a warp's two frames of a tile are neighbours.  A work unit is `tiles_per_unit` tiles of one capture;
 * like in spectrum.cuh the unit's four warps are added in fixed order into one 1024-float partial per unit and
 * k_spectrum_finalize adds the units in order: the bits of a capture's spectrum depend on its length only.
 *
 * the fifth warp to the other four: 240 / 40):
 *   warps 0..3  the FFT warps, warp w = TMEM lanes 32 w ...  Per half-tile s of the CTA's stream:
 *                 wait for MMA(s)                  [mbarrier, committed two half-tiles ago]
 *                 D -> 64 registers                [tcgen05.ld, asynchronous]
 *                 rows of half-tile s + 2 -> A     [32 LDS.U16 out of the frame TMA dropped into shared memory, PRMT,
 *                                                   HADD2, tcgen05.st, asynchronous]
 *                 twiddle, window, transpose store of half-tile s
 *                 tcgen05.wait::st, arrive on "A full" (the store has long finished: no stall)
 *                 transpose load, FFT-32, |X|^2
 *   warp 4      one lane: bulk copies (cp.async.bulk, 2048 contiguous bytes = one frame per FFT warp) two half-tiles
 *               ahead into a three-slot ring; when the four FFT warps have arrived on "A full" it issues the eight
 *               tcgen05.mma of the half-tile and commits to "D full".  It also draws the units from the counter and
 *               publishes them through a sequence-tagged ring.
 * No CTA-wide barrier in the loop; at the end of a unit the FFT warps meet at an mbarrier (the fixed-order reduction).
 * Every wait is bounded (clock64): a protocol error sets *error and ends the kernel.
 */
#ifndef B200_SPEC_TC_SPIKE_CUH
#define B200_SPEC_TC_SPIKE_CUH

#include "../stm32f7-rtlsdr_b200/csrc/plan.h" /* spectrum.cuh, tma.cuh, wbfm_tc.cuh (mbarrier / descriptor / tcgen05 helpers) */

#define B200_STC_CONSUMERS 128
#define B200_STC_THREADS 160
#define B200_STC_N 64                               /* (c, re/im)                                               */
#define B200_STC_B_BOX (B200_STC_N * 64)            /* 64 rows x 64 bytes of K (32 fp16), 64B swizzle           */
#define B200_STC_B_PART (2 * B200_STC_B_BOX)        /* K = 64 fp16 = two boxes                                  */
#define B200_STC_B_BYTES (2 * B200_STC_B_PART)      /* hi part, lo part: 16 KiB                                 */
#define B200_STC_D_COLS 64
#define B200_STC_A_COLS 32
#define B200_STC_TMEM_COLS 256                      /* allocated (power of two): D even 0, D odd 64, A even 128, A odd 160 */
#define B200_STC_COL_D(p) ((uint32_t)(p) * B200_STC_D_COLS)
#define B200_STC_COL_A(p) (128u + (uint32_t)(p) * B200_STC_A_COLS)
#define B200_STC_FRAMES_PER_TILE 8
#define B200_STC_RING 8                             /* unit hand-out ring (sequence-tagged)                     */
#define B200_STC_RAW_SLOTS 3                        /* half-tiles of raw frames in flight                       */
#define B200_STC_FRAME_BYTES 2048
#define B200_STC_RAW_SLOT_BYTES (4 * B200_STC_FRAME_BYTES)
#define B200_STC_LOOKAHEAD 2                        /* the producer copies this many half-tiles ahead of the MMA it issues */

#define B200_STC_SM_B 0
#define B200_STC_SM_XP (B200_STC_SM_B + B200_STC_B_BYTES)                          /* c2 [4][32][XP]            */
#define B200_STC_SM_RAW (B200_STC_SM_XP + B200_SPEC_WARPS * 32 * B200_SPEC_XP * 8) /* u8 [slots][4][2048]       */
#define B200_STC_SM_RED (B200_STC_SM_RAW + B200_STC_RAW_SLOTS * B200_STC_RAW_SLOT_BYTES) /* float [2][4][1024]  */
#define B200_STC_SM_BAR (B200_STC_SM_RED + 2 * 4 * 1024 * 4)                       /* u64 [12]                  */
#define B200_STC_SM_MISC (B200_STC_SM_BAR + 12 * 8) /* u32: [0] TMEM base, [1] abort, [2..3] conversion constants */
#define B200_STC_SM_UNITS (B200_STC_SM_MISC + 32)                                  /* u32 [RING][2] (tag, unit) */
#define B200_STC_SMEM_BYTES (B200_STC_SM_UNITS + B200_STC_RING * 8)
/* mbarriers */
#define B200_STC_BAR_RAW_FULL 0   /* [3] producer's expect_tx + the bytes                */
#define B200_STC_BAR_RAW_EMPTY 3  /* [3] one arrival per FFT warp                        */
#define B200_STC_BAR_A_FULL 6     /* [2] one arrival per FFT warp                        */
#define B200_STC_BAR_D_FULL 8     /* [2] tcgen05.commit                                  */
#define B200_STC_BAR_RED 10       /* [2] one arrival per FFT warp                        */

struct SpectrumTcParams {
    const uint8_t *iq;          /* capture c starts at iq + c * capture_stride (16-byte aligned, stride too)    */
    uint64_t capture_stride;    /* bytes                                                                        */
    uint32_t frames;            /* frames per capture                                                           */
    uint32_t tiles_per_unit;    /* a unit = 8 x tiles_per_unit consecutive frames of one capture                */
    uint32_t units_per_capture;
    uint32_t total_units;
    const float2 *twiddle;      /* 1024 entries e^{-2 pi i m / 1024}                                            */
    const uint8_t *b_image;     /* B200_STC_B_BYTES: the B operand as it lies in shared memory                  */
    float *partials;            /* [capture][units_per_capture][1024]                                           */
    uint32_t *unit_counter;     /* as in SpectrumParams                                                         */
    uint32_t *error;            /* device word, 0 = ok                                                          */
    float ema_log2_decay, ema_beta;
    float *dbg_y0;              /* optional [128][64]: raw accumulators of the first half-tile of unit 0 (tests) */
    unsigned long long *dbg_prof; /* optional [8]: cycles warp 0 of CTA 0 spent in each phase (-DB200_STC_PROFILE)  */
    uint32_t dbg_flags;         /* timing experiments only (results wrong): 1 no frame arithmetic, 2 no conversion + TMEM store,
                                   4 no MMAs (commit only), 8 no bulk copies, 16 no TMEM loads                   */
};

#if defined(__CUDACC__) && !defined(B200_EMULATED)

/* D = F32 (1 at [4,6)), A = B = F16 (0 at [7,10), [10,13)), both K-major, N >> 3 at [17,23), M >> 4 at [24,29) */
#define B200_STC_IDESC ((1u << 4) | ((uint32_t)(B200_STC_N >> 3) << 17) | ((128u >> 4) << 24))

B200_DEV void b200_stc_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(B200_STC_IDESC), "r"(accumulate)
        : "memory");
}
B200_DEV void b200_stc_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
/* raw word (I | Q << 8 | ...) -> fp16 pair (I - 128, Q - 128): PRMT builds (0x6400 | I, 0x6400 | Q) = (1024 + I, 1024 + Q),
 * one packed half add of -1152 leaves the integers exactly */
B200_DEV uint32_t b200_stc_cvt(uint32_t raw, uint32_t magic, uint32_t bias)
{
    uint32_t h = __byte_perm(raw, magic, 0x5140), r;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(bias));
    return r;
}

/* one frame of one lane, first half: d = the 64 accumulator words Y0'_b[c] (re, im), b = lane.  Twiddle, window, and the
 * transpose store: row c, column b of the warp's tile */
B200_DEV void b200_stc_frame_a(const uint32_t (&d)[64], const float (&twr)[32], const float (&twi)[32], float wr, float wi,
                               c2 *s_xp, int lane)
{
    c2 z[32];
    /* the half-sample offset lives in column 0: + 16 (1 + i); tw[0] = 1/4 */
    z[0] = c2_scale(c2_add(c2_make(__uint_as_float(d[0]), __uint_as_float(d[1])), c2_make(16.0f, 16.0f)), 0.25f);
#pragma unroll
    for (int c = 1; c < 32; ++c) z[c] = c2_cmul(c2_make(__uint_as_float(d[2 * c]), __uint_as_float(d[2 * c + 1])), twr[c], twi[c]);
    const c2 zm = c2_cmul(z[31], wr, -wi); /* Z[-1] = W32^{-b} Z[31] */
    const c2 zp = c2_cmul(z[0], wr, wi);   /* Z[32] = W32^{ b} Z[0]  */
    __syncwarp(); /* the previous frame's transposed loads are done */
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const c2 nb = c2_add(c == 0 ? zm : z[c == 0 ? 0 : c - 1], c == 31 ? zp : z[c == 31 ? 31 : c + 1]);
        s_xp[c * B200_SPEC_XP + lane] = c2_two_a_minus(z[c], nb);
    }
}
/* second half: transposed load (this lane = column c of the spectrum), second FFT-32, |X|^2 into acc[d] (bin lane + 32 d) */
template <bool EMA>
B200_DEV void b200_stc_frame_b(const c2 *s_xp, int lane, float (&acc)[32], float wgt)
{
    __syncwarp();
    c2 v[32];
#pragma unroll
    for (int t = 0; t < 32; t += 2) {
        const float4 q = *reinterpret_cast<const float4 *>(s_xp + lane * B200_SPEC_XP + t);
        v[b200_bitrev5(t)] = c2_make(q.x, q.y);
        v[b200_bitrev5(t + 1)] = c2_make(q.z, q.w);
    }
    b200_fft32(v); /* v[d] = X[lane + 32 d] */
    if (EMA) {
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) acc[k2] = fmaf(c2_norm_acc(v[k2], 0.0f), wgt, acc[k2]);
    } else {
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) acc[k2] = c2_norm_acc(v[k2], acc[k2]);
    }
}

/* The CTA's stream of half-tiles: units in hand-out order, half-tiles of a unit in order.  The producer lane draws the
 * units and publishes them; everybody else reads them from the ring. */
struct StcWalker {
    uint32_t q, unit, h, H, m0, mend, capture;
    bool alive;
};
struct StcSlot {
    uint32_t unit, m, flags; /* flags: 1 exists, 2 this warp's frame m is inside the capture, 4 last half-tile of its unit */
};
B200_DEV void b200_stc_enter(const SpectrumTcParams &p, StcWalker &w, uint32_t unit)
{
    w.unit = unit;
    w.alive = unit < p.total_units;
    if (!w.alive) return;
    const uint32_t unit_frames = p.tiles_per_unit * B200_STC_FRAMES_PER_TILE;
    w.capture = unit / p.units_per_capture;
    w.m0 = (unit - w.capture * p.units_per_capture) * unit_frames;
    w.mend = w.m0 + unit_frames;
    if (w.mend > p.frames) w.mend = p.frames;
    w.H = 2u * ((w.mend - w.m0 + B200_STC_FRAMES_PER_TILE - 1u) / B200_STC_FRAMES_PER_TILE);
    w.h = 0;
}
/* frame of FFT warp `warp` in the walker's current half-tile */
B200_DEV uint32_t b200_stc_frame_of(const StcWalker &w, int warp) { return w.m0 + B200_STC_FRAMES_PER_TILE * (w.h >> 1) + 2u * (uint32_t)warp + (w.h & 1u); }
B200_DEV bool b200_stc_ring_read(volatile uint32_t *s_units, uint32_t q, volatile uint32_t *abort_flag, uint32_t &unit)
{
    volatile uint32_t *e = s_units + 2 * (q % B200_STC_RING);
    const long long t0 = clock64();
    while (e[0] != q) {
        if (*abort_flag || clock64() - t0 > (1ll << 28)) return false;
    }
    unit = e[1];
    return true;
}

template <bool EMA>
__global__ void __maxnreg__(192) k_spectrum_tc(SpectrumTcParams p)
{
    B200_DYN_SMEM(smem);
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + B200_STC_SM_BAR);
    uint32_t *s_misc = reinterpret_cast<uint32_t *>(smem + B200_STC_SM_MISC);
    volatile uint32_t *s_abort = s_misc + 1;
    volatile uint32_t *s_units = reinterpret_cast<uint32_t *>(smem + B200_STC_SM_UNITS);

    if (tid == 0) {
        for (int i = 0; i < B200_STC_RAW_SLOTS; ++i) {
            b200_tc_bar_init(s_bar + B200_STC_BAR_RAW_FULL + i, 1);
            b200_tc_bar_init(s_bar + B200_STC_BAR_RAW_EMPTY + i, 4);
        }
        for (int i = 0; i < 2; ++i) {
            b200_tc_bar_init(s_bar + B200_STC_BAR_A_FULL + i, 4);
            b200_tc_bar_init(s_bar + B200_STC_BAR_D_FULL + i, 1);
            b200_tc_bar_init(s_bar + B200_STC_BAR_RED + i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_misc[1] = 0u;
        s_misc[2] = 0x64646464u;
        s_misc[3] = 0xE480E480u; /* -1152 as fp16, twice */
        for (int i = 0; i < B200_STC_RING; ++i) s_units[2 * i] = 0xffffffffu;
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(b200_tc_smem(s_misc)), "r"((uint32_t)B200_STC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.b_image);
        uint4 *dst = reinterpret_cast<uint4 *>(smem + B200_STC_SM_B);
        for (int i = tid; i < B200_STC_B_BYTES / 16; i += B200_STC_THREADS) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_misc[0];

    if (warp == 4) {
        /* ===== producer + MMA issuer (one lane) ===== */
        if (lane == 0) {
            const uint32_t b_base = b200_tc_smem(smem + B200_STC_SM_B);
            uint8_t *s_raw = smem + B200_STC_SM_RAW;
            StcWalker w;
            w.q = 0;
            uint32_t fail = 0, copied = 0;
            /* publish + enter the first unit */
            auto publish = [&](uint32_t q, uint32_t unit) {
                volatile uint32_t *e = s_units + 2 * (q % B200_STC_RING);
                e[1] = unit;
                __threadfence_block();
                e[0] = q;
            };
            publish(0u, blockIdx.x);
            b200_stc_enter(p, w, blockIdx.x);
            auto copy_next = [&]() { /* bulk copies of the walker's half-tile, then advance it */
                const uint32_t slot = copied % B200_STC_RAW_SLOTS, k = copied / B200_STC_RAW_SLOTS;
                if (!b200_tc_wait(s_bar + B200_STC_BAR_RAW_EMPTY + slot, (k & 1u) ^ 1u, s_abort)) { fail = 5; return; }
                const uint8_t *cap = p.iq + (uint64_t)w.capture * p.capture_stride;
                uint32_t nvalid = 0;
#pragma unroll
                for (int f = 0; f < 4; ++f) nvalid += b200_stc_frame_of(w, f) < w.mend ? 1u : 0u;
                if (p.dbg_flags & 8u) nvalid = 0;
                if (nvalid) b200_mbar_expect_tx(s_bar + B200_STC_BAR_RAW_FULL + slot, nvalid * B200_STC_FRAME_BYTES);
                else b200_tc_arrive(s_bar + B200_STC_BAR_RAW_FULL + slot);
                if (nvalid) {
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const uint32_t m = b200_stc_frame_of(w, f);
                        if (m < w.mend)
                            b200_tma_load_1d(s_raw + slot * B200_STC_RAW_SLOT_BYTES + f * B200_STC_FRAME_BYTES, cap + (uint64_t)m * 1024u,
                                             B200_STC_FRAME_BYTES, s_bar + B200_STC_BAR_RAW_FULL + slot);
                    }
                }
                ++copied;
                if (++w.h == w.H) {
                    ++w.q;
                    const uint32_t next = gridDim.x + atomicAdd(p.unit_counter, 1u);
                    publish(w.q, next);
                    b200_stc_enter(p, w, next);
                }
            };
            for (int i = 0; i < B200_STC_LOOKAHEAD && w.alive && !fail; ++i) copy_next();
            for (uint32_t j = 0; !fail; ++j) {
                if (w.alive) copy_next();
                if (j >= copied || fail) break;
                const uint32_t par = j & 1u;
                if (!b200_tc_wait(s_bar + B200_STC_BAR_A_FULL + par, (j >> 1) & 1u, s_abort)) { fail = 6; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (p.dbg_flags & 4u) break;
                    const uint32_t boff = (uint32_t)((ks >> 1) * B200_STC_B_BOX + (ks & 1) * 32);
                    b200_stc_mma(tmem + B200_STC_COL_D(par), tmem + B200_STC_COL_A(par) + 8u * ks, b200_tc_desc(b_base + boff), ks > 0 ? 1u : 0u);
                    b200_stc_mma(tmem + B200_STC_COL_D(par), tmem + B200_STC_COL_A(par) + 8u * ks, b200_tc_desc(b_base + B200_STC_B_PART + boff), 1u);
                }
                b200_tc_commit(s_bar + B200_STC_BAR_D_FULL + par);
            }
            if (fail) { *s_abort = 1u; atomicCAS(p.error, 0u, fail); }
        }
        __syncwarp();
    } else {
        /* ===== FFT warps ===== */
        c2 *s_xp = reinterpret_cast<c2 *>(smem + B200_STC_SM_XP) + warp * (32 * B200_SPEC_XP);
        float *s_red = reinterpret_cast<float *>(smem + B200_STC_SM_RED);
        const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
        /* this lane's constants: 1/4 W1024^{b c}, c = 0..31, and W32^{b} */
        float twr[32], twi[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float2 w = __ldg(p.twiddle + ((lane * c) & 1023));
            twr[c] = 0.25f * w.x;
            twi[c] = 0.25f * w.y;
        }
        const float2 wrap = __ldg(p.twiddle + 32 * lane);
        /* conversion constants out of shared memory, so each stays in one ordinary register (cplx2.cuh) */
        const uint32_t cv_magic = s_misc[2], cv_bias = s_misc[3];
        bool ok = true;
        StcWalker w;
        w.q = 0;
        {
            uint32_t u = 0;
            ok = b200_stc_ring_read(s_units, 0u, s_abort, u);
            b200_stc_enter(p, w, ok ? u : 0xffffffffu);
        }
        uint32_t walked = 0; /* half-tiles handed out by next_slot */
        auto next_slot = [&]() -> StcSlot {
            StcSlot s;
            s.unit = 0; s.m = 0; s.flags = 0;
            if (!w.alive || !ok) return s;
            s.unit = w.unit;
            s.m = b200_stc_frame_of(w, warp);
            s.flags = 1u | (s.m < w.mend ? 2u : 0u) | (w.h + 1u == w.H ? 4u : 0u);
            ++walked;
            if (++w.h == w.H) {
                ++w.q;
                uint32_t u = 0;
                ok = b200_stc_ring_read(s_units, w.q, s_abort, u);
                b200_stc_enter(p, w, ok ? u : 0xffffffffu);
            }
            return s;
        };
        /* rows of half-tile number idx (slot sl of the stream) -> A operand: shared memory -> fp16 pairs -> tcgen05.st (asynchronous) */
        auto stage_begin = [&](const StcSlot &sl, uint32_t idx) {
            const uint32_t slot = idx % B200_STC_RAW_SLOTS, k = idx / B200_STC_RAW_SLOTS;
            if (!b200_tc_wait(s_bar + B200_STC_BAR_RAW_FULL + slot, k & 1u, s_abort)) { ok = false; return; }
            if ((sl.flags & 2u) && !(p.dbg_flags & 2u)) {
                const unsigned short *src = reinterpret_cast<const unsigned short *>(smem + B200_STC_SM_RAW + slot * B200_STC_RAW_SLOT_BYTES + warp * B200_STC_FRAME_BYTES) + lane;
                uint32_t h[32];
#pragma unroll
                for (int a = 0; a < 32; ++a) h[a] = b200_stc_cvt((uint32_t)src[32 * a], cv_magic, cv_bias);
                b200_stc_st32(tlane + B200_STC_COL_A(idx & 1u), h);
            }
            __syncwarp();
            if (lane == 0) b200_tc_arrive(s_bar + B200_STC_BAR_RAW_EMPTY + slot);
        };
        auto stage_end = [&](uint32_t idx) {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) b200_tc_arrive(s_bar + B200_STC_BAR_A_FULL + (idx & 1u));
        };

        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
        uint32_t units_done = 0;
#ifdef B200_STC_PROFILE
        long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tp = clock64();
#define STC_TICK(k) do { const long long t_ = clock64(); prof[k] += t_ - tp; tp = t_; } while (0)
#else
#define STC_TICK(k) do { } while (0)
#endif

        StcSlot cur = next_slot(), n1 = next_slot();
        if ((cur.flags & 1u) && ok) { stage_begin(cur, 0u); stage_end(0u); }
        if ((n1.flags & 1u) && ok) { stage_begin(n1, 1u); stage_end(1u); }

        for (uint32_t i = 0; (cur.flags & 1u) && ok; ++i) {
            const uint32_t par = i & 1u;
            const StcSlot n2 = next_slot();
            if (!ok) break;
            STC_TICK(0);
            if (!b200_tc_wait(s_bar + B200_STC_BAR_D_FULL + par, (i >> 1) & 1u, s_abort)) { ok = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            STC_TICK(1);
            const bool work = (cur.flags & 2u) != 0u;
            /* A of this parity is free (MMA(i) is done): the rows of half-tile i + 2 first -- their registers are gone again
             * before the accumulators arrive */
            if (n2.flags & 1u) {
                stage_begin(n2, i + 2u);
                if (!ok) break;
            }
            STC_TICK(2);
            uint32_t d[64];
            if (work && !(p.dbg_flags & 16u)) {
                uint32_t d_lo[32], d_hi[32];
                b200_tc_ld32(tlane + B200_STC_COL_D(par), d_lo);
                b200_tc_ld32(tlane + B200_STC_COL_D(par) + 32u, d_hi);
#pragma unroll
                for (int c = 0; c < 32; ++c) { d[c] = d_lo[c]; d[32 + c] = d_hi[c]; }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            STC_TICK(3);
            if (work && p.dbg_y0 && i == 0 && cur.unit == 0) {
#pragma unroll
                for (int c = 0; c < 64; ++c) p.dbg_y0[tid * 64 + c] = __uint_as_float(d[c]);
            }
            if (work && !(p.dbg_flags & 1u)) b200_stc_frame_a(d, twr, twi, wrap.x, wrap.y, s_xp, lane);
            STC_TICK(4);
            if (n2.flags & 1u) stage_end(i + 2u);
            STC_TICK(5);
            if (work && !(p.dbg_flags & 1u)) {
                float wgt = 0.0f;
                if (EMA) wgt = p.ema_beta * exp2f((float)(p.frames - 1u - cur.m) * p.ema_log2_decay);
                b200_stc_frame_b<EMA>(s_xp, lane, acc, wgt);
            }
            STC_TICK(6);
            if (cur.flags & 4u) {
                /* end of a unit: fixed-order sum over the four warps, one partial per unit.  Two buffers: the warps are never
                 * more than two half-tiles (= at most one unit) apart */
                const uint32_t buf = units_done & 1u;
                float *mine = s_red + (buf * 4 + warp) * 1024;
#pragma unroll
                for (int k2 = 0; k2 < 32; ++k2) { mine[k2 * 32 + lane] = acc[k2]; acc[k2] = 0.0f; }
                __syncwarp();
                if (lane == 0) b200_tc_arrive(s_bar + B200_STC_BAR_RED + buf);
                if (!b200_tc_wait(s_bar + B200_STC_BAR_RED + buf, (units_done >> 1) & 1u, s_abort)) { ok = false; break; }
                float *out = p.partials + (uint64_t)cur.unit * 1024u;
#pragma unroll
                for (int k = 0; k < 1024; k += B200_STC_CONSUMERS) {
                    float s = 0.0f;
#pragma unroll
                    for (int ww = 0; ww < 4; ++ww) s += s_red[(buf * 4 + ww) * 1024 + k + tid];
                    out[k + tid] = s;
                }
                ++units_done;
            }
            cur = n1;
            n1 = n2;
            STC_TICK(7);
        }
#ifdef B200_STC_PROFILE
        if (p.dbg_prof && blockIdx.x == 0 && tid == 0)
            for (int k = 0; k < 8; ++k) p.dbg_prof[k] = (unsigned long long)prof[k];
#endif
        if (!ok && lane == 0) { *s_abort = 1u; atomicCAS(p.error, 0u, 1u + (uint32_t)warp); }
    }

    /* the last CTA out leaves the hand-out counter at zero for the next launch */
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const bool last = atomicInc(p.unit_counter + 1, gridDim.x - 1u) == gridDim.x - 1u;
        if (last) p.unit_counter[0] = 0u;
    }
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)B200_STC_TMEM_COLS) : "memory");
}

#endif /* device */

namespace b200 {
/* A unit = tiles_per_unit tiles of 8 consecutive frames; like frames_per_warp of the FP32 engine (two of a warp's frames
 * per tile) it follows from the capture LENGTH only. */
struct SpectrumTcPlan {
    uint32_t frames, tiles_per_unit, units_per_capture, grid;
    uint64_t total_units;
};
inline SpectrumTcPlan plan_spectrum_tc(uint64_t len_bytes, uint32_t n_captures, uint32_t sm_count)
{
    SpectrumTcPlan pl{};
    pl.frames = (uint32_t)spectrum_frames(len_bytes);
    if (pl.frames == 0) return pl;
    pl.tiles_per_unit = (spectrum_frames_per_warp(pl.frames) + 1u) / 2u;
    pl.units_per_capture = (uint32_t)ceil_div(pl.frames, (uint64_t)pl.tiles_per_unit * B200_STC_FRAMES_PER_TILE);
    pl.total_units = (uint64_t)pl.units_per_capture * n_captures;
    const uint64_t slots = (uint64_t)sm_count * 2u;
    pl.grid = (uint32_t)(pl.total_units < slots ? pl.total_units : slots);
    return pl;
}
/* IEEE binary16 <-> double, round to nearest even (values of magnitude <= 2 only: no overflow handling needed) */
inline uint16_t f16_from_double(double x)
{
    const uint16_t sign = std::signbit(x) ? 0x8000u : 0u;
    const double ax = std::fabs(x);
    if (ax == 0.0) return sign;
    int e;
    (void)std::frexp(ax, &e); /* ax = m 2^e, m in [0.5, 1) */
    int E = e - 1;
    if (E < -14) return (uint16_t)(sign | (uint16_t)std::nearbyint(std::ldexp(ax, 24))); /* subnormal (1024 = smallest normal) */
    double mant = std::nearbyint((std::ldexp(ax, -E) - 1.0) * 1024.0);
    if (mant == 1024.0) { mant = 0.0; ++E; }
    return (uint16_t)(sign | (uint16_t)((E + 15) << 10) | (uint16_t)mant);
}
inline double f16_to_double(uint16_t h)
{
    const int E = (h >> 10) & 31, m = h & 1023;
    const double v = E == 0 ? std::ldexp((double)m, -24) : std::ldexp(1.0 + m / 1024.0, E - 15);
    return (h & 0x8000u) ? -v : v;
}
/* The DFT-32 matrix of pass 1 as the B operand: rows n = 2 c + (re = 0 | im = 1), K index k = 2 a + (I = 0 | Q = 1);
 * (I + i Q)(cos t - i sin t), t = 2 pi a c / 32: re = I cos t + Q sin t, im = Q cos t - I sin t.  Two parts, fp16 hi and
 * fp16 lo of the float64 value (22 bits), each two 64-byte-swizzled boxes of 32 K elements. */
inline void fill_spectrum_tc_image(uint8_t *image)
{
    std::memset(image, 0, B200_STC_B_BYTES);
    for (int c = 0; c < 32; ++c)
        for (int a = 0; a < 32; ++a) {
            const int r = (a * c) & 31;
            const double cs = (r == 8 || r == 24) ? 0.0 : std::cos(2.0 * kPi * r / 32.0);
            const double sn = (r == 0 || r == 16) ? 0.0 : std::sin(2.0 * kPi * r / 32.0);
            const double val[2][2] = {{cs, sn}, {-sn, cs}}; /* [re/im][I/Q] */
            for (int ri = 0; ri < 2; ++ri)
                for (int iq = 0; iq < 2; ++iq) {
                    const int n = 2 * c + ri, k = 2 * a + iq;
                    const uint16_t hi = f16_from_double(val[ri][iq]);
                    const uint16_t lo = f16_from_double(val[ri][iq] - f16_to_double(hi));
                    const uint32_t off = B200_TC_OP_OFF(B200_STC_N, n, 2 * k);
                    std::memcpy(image + off, &hi, 2);
                    std::memcpy(image + B200_STC_B_PART + off, &lo, 2);
                }
        }
}
} // namespace b200

#endif
