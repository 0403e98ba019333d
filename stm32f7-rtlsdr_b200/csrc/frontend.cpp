/*
 * frontend.cpp -- RTL2832 / E4000 parameter math behind include/b200sdr_frontend.h.
 * Host-only, integer / double arithmetic; every function names the reference routine it restates
 * and is pinned bit-for-bit against that routine by tests/test_frontend_parity.py (oracle A).
 */
#include "../../include/b200sdr_frontend.h"

extern "C" {

/* RTLSDR_set_sample_rate, RTL/Src/usbh_rtlsdr.c:683-689: ratio = xtal * 2^22 / rate in double,
 * truncated to 32 bits, low two bits and the top nibble cleared; bit 27 is mirrored into bit 28 by
 * the hardware; the achieved rate is xtal * 2^22 / that. */
int32_t b200sdr_rtl_resampler(uint32_t samp_rate, uint32_t xtal_hz, b200sdr_rtl_rate *out)
{
    if (!out || samp_rate == 0) return B200SDR_FAIL;
    const double scaled_xtal = (double)xtal_hz * 4194304.0;
    uint32_t ratio = (uint32_t)(scaled_xtal / (double)samp_rate);
    ratio &= 0x0FFFFFFCu;
    const uint32_t applied = ratio | ((ratio & 0x08000000u) << 1);
    out->rsamp_ratio = ratio;
    out->real_rsamp_ratio = applied;
    out->real_rate = scaled_xtal / (double)applied;
    const bool unsupported = samp_rate <= 225000u || samp_rate > 3200000u || (samp_rate > 300000u && samp_rate <= 900000u);
    return unsupported ? B200SDR_NOT_SUPPORTED : B200SDR_OK;
}

/* RTL/Inc/usbh_rtlsdr.h:340-345 (the DAB/FM coefficients of the vendor driver) */
void b200sdr_rtl_default_fir(int32_t coeff[16])
{
    static const int32_t table[16] = {-54, -36, -41, -40, -32, -14, 14, 53, 101, 156, 215, 273, 327, 372, 404, 421};
    for (int i = 0; i < 16; ++i) coeff[i] = table[i];
}

/* RTLSDR_set_fir, RTL/Src/usbh_rtlsdr.c:552-572: eight int8, then eight int12 packed two per three
 * bytes: [hi 8 of a] [lo 4 of a | hi 4 of b] [lo 8 of b]. */
int32_t b200sdr_rtl_fir_pack(const int32_t coeff[16], uint8_t out20[20])
{
    if (!coeff || !out20) return B200SDR_FAIL;
    bool in_range = true;
    for (int i = 0; i < 8; ++i) {
        in_range = in_range && coeff[i] >= -128 && coeff[i] <= 127;
        out20[i] = (uint8_t)coeff[i];
    }
    for (int pair = 0; pair < 4; ++pair) {
        const int32_t a = coeff[8 + 2 * pair], b = coeff[9 + 2 * pair];
        in_range = in_range && a >= -2048 && a <= 2047 && b >= -2048 && b <= 2047;
        uint8_t *dst = out20 + 8 + 3 * pair;
        dst[0] = (uint8_t)(a >> 4);
        dst[1] = (uint8_t)((a << 4) | ((b >> 8) & 0x0F));
        dst[2] = (uint8_t)b;
    }
    return in_range ? B200SDR_OK : B200SDR_NOT_SUPPORTED;
}

/* E4K_compute_pll_params, RTL/Src/tuner_e4k.c:689-737 with its band table :301-312 and
 * compute_fvco / compute_flo :338-361.  Fvco = Fosc (Z + X / 65536), Flo = Fvco / R.  The
 * reference stores Z in 8 and X in 16 bits before recomputing Flo; the same narrowing is applied. */
uint32_t b200sdr_e4k_pll_params(uint32_t fosc, uint32_t intended_flo, b200sdr_e4k_pll *out)
{
    struct Band { uint32_t below_hz; uint8_t synth7, divider; };
    static const Band bands[] = {
        {72400000u, 0x0F, 48}, {81200000u, 0x0E, 40}, {108300000u, 0x0D, 32}, {162500000u, 0x0C, 24},
        {216600000u, 0x0B, 16}, {325000000u, 0x0A, 12}, {350000000u, 0x09, 8}, {432000000u, 0x03, 8},
        {667000000u, 0x02, 6},  {1200000000u, 0x01, 4},
    };
    if (!out) return 0;
    out->r_idx = 0;
    if (fosc < 16000000u || fosc > 30000000u) return 0;
    uint8_t divider = 2, threephase = 0;
    for (const Band &b : bands) {
        if (intended_flo < b.below_hz) {
            threephase = (b.synth7 & 0x08) ? 1 : 0;
            out->r_idx = b.synth7;
            divider = b.divider;
            break;
        }
    }
    const uint64_t want_vco = (uint64_t)intended_flo * divider;
    const uint64_t whole = want_vco / fosc;
    const uint64_t rest = want_vco - (uint64_t)fosc * whole;
    const uint32_t frac = (uint32_t)((rest * 65536u) / fosc);
    const uint8_t z8 = (uint8_t)whole;
    const uint16_t x16 = (uint16_t)frac;
    const uint64_t vco = (uint64_t)fosc * z8 + ((uint64_t)fosc * x16) / 65536u;
    const uint32_t flo = (uint32_t)(vco / divider);
    out->fosc = fosc;
    out->flo = flo;
    out->intended_flo = intended_flo;
    out->r = divider;
    out->threephase = threephase;
    out->x = x16;
    out->z = z8;
    return flo;
}

} /* extern "C" */
