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
    /* the rate check comes FIRST (the reference checks a compile-time constant, :677-681): below 28 126 S/s at
     * 28.8 MHz the quotient does not fit 32 bits and the conversion would be undefined; *out stays zeroed */
    out->rsamp_ratio = out->real_rsamp_ratio = 0;
    out->real_rate = 0.0;
    if (samp_rate <= 225000u || samp_rate > 3200000u || (samp_rate > 300000u && samp_rate <= 900000u)) return B200SDR_NOT_SUPPORTED;
    const double scaled_xtal = (double)xtal_hz * 4194304.0;
    const double quotient = scaled_xtal / (double)samp_rate;
    if (!(quotient < 4294967296.0)) return B200SDR_NOT_SUPPORTED; /* absurd crystal: same guard */
    uint32_t ratio = (uint32_t)quotient;
    ratio &= 0x0FFFFFFCu;
    const uint32_t applied = ratio | ((ratio & 0x08000000u) << 1);
    out->rsamp_ratio = ratio;
    out->real_rsamp_ratio = applied;
    out->real_rate = scaled_xtal / (double)applied;
    return B200SDR_OK;
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

/* USBH_RTLSDR_ClassRequest, RTL/Src/usbh_rtlsdr.c:809-945, as a list of control transfers.  The emitters below
 * restate the setup-packet encodings of RTLSDR_write_reg (:262-300), RTLSDR_demod_write_reg (:470-521, including
 * the dummy read that follows every demod write, RTLSDR_demod_read_reg :445-468) and RTLSDR_i2c_read_reg
 * (:347-398: one-byte write of the register number, one-byte read, both through block IICB). */
namespace {
struct SeqWriter {
    b200sdr_ctl_xfer *out;
    uint32_t capacity, n;
    uint8_t step;
    void emit(uint8_t type, uint16_t value, uint16_t index, uint16_t len, uint16_t val)
    {
        if (n < capacity) {
            b200sdr_ctl_xfer x{};
            x.bmRequestType = type;
            x.wValue = value;
            x.wIndex = index;
            x.wLength = len;
            if (type == 0x40) { /* 16-bit values travel high byte first */
                x.data[0] = (uint8_t)(len == 1 ? val & 0xFF : val >> 8);
                x.data[1] = (uint8_t)(len == 1 ? 0 : val & 0xFF);
            }
            x.step = step;
            out[n] = x;
        }
        ++n;
    }
    void write_reg(uint8_t block, uint16_t addr, uint16_t val, uint8_t len) { emit(0x40, addr, (uint16_t)((block << 8) | 0x10), len, val); }
    void demod_write(uint8_t page, uint16_t addr, uint16_t val, uint8_t len)
    {
        emit(0x40, (uint16_t)((addr << 8) | 0x20), (uint16_t)(0x10 | page), len, val);
        emit(0xC0, (uint16_t)((0x01 << 8) | 0x20), 0x0a, 1, 0);
    }
    void i2c_read_reg(uint8_t chip, uint8_t reg)
    {
        emit(0x40, chip, (uint16_t)((6 << 8) | 0x10), 1, reg);
        emit(0xC0, chip, (uint16_t)(6 << 8), 1, 0);
    }
};
} // namespace

int32_t b200sdr_rtl_init_sequence(uint32_t samp_rate, uint32_t xtal_hz, const int32_t *fir16, uint32_t flags,
                                  b200sdr_ctl_xfer *out, uint32_t capacity, uint32_t *n_out)
{
    if (!n_out || (!out && capacity) || samp_rate == 0) return B200SDR_FAIL;
    enum : uint8_t { USBB = 1, SYSB = 2 };                                  /* RTL/Inc/usbh_rtlsdr.h:319-327 */
    enum : uint16_t { USB_SYSCTL = 0x2000, USB_EPA_CTL = 0x2148, USB_EPA_MAXPKT = 0x2158, DEMOD_CTL = 0x3000, DEMOD_CTL_1 = 0x300b };
    int32_t coeff[16];
    if (fir16) for (int i = 0; i < 16; ++i) coeff[i] = fir16[i];
    else b200sdr_rtl_default_fir(coeff);
    uint8_t fir[20];
    const int32_t fir_rc = b200sdr_rtl_fir_pack(coeff, fir);
    b200sdr_rtl_rate rate{};
    const int32_t rate_rc = b200sdr_rtl_resampler(samp_rate, xtal_hz, &rate);
    if (fir_rc != B200SDR_OK || rate_rc != B200SDR_OK) { /* nothing is emitted for parameters the chip cannot take */
        *n_out = 0;
        return B200SDR_NOT_SUPPORTED;
    }

    SeqWriter w{out, capacity, 0, 0};
    w.step = 0;  w.write_reg(USBB, USB_SYSCTL, 0x09, 1);            /* dummy write */
    w.step = 1;  w.write_reg(USBB, USB_SYSCTL, 0x09, 1);            /* initialise the USB block */
    w.step = 2;  w.write_reg(USBB, USB_EPA_MAXPKT, 0x0002, 2);
    w.step = 3;  w.write_reg(USBB, USB_EPA_CTL, 0x1002, 2);
    w.step = 4;  w.write_reg(SYSB, DEMOD_CTL_1, 0x22, 1);           /* power the demodulator on */
    w.step = 5;  w.write_reg(SYSB, DEMOD_CTL, 0xe8, 1);
    w.step = 6;  w.demod_write(1, 0x01, 0x14, 1);                   /* soft reset on, off */
    w.step = 7;  w.demod_write(1, 0x01, 0x10, 1);
    w.step = 8;  w.demod_write(1, 0x15, 0x00, 1);                   /* no spectrum inversion / adjacent channel rejection */
    w.step = 9;  w.demod_write(1, 0x16, 0x0000, 2);
    for (uint8_t i = 0; i < 6; ++i) { w.step = (uint8_t)(10 + i); w.demod_write(1, (uint16_t)(0x16 + i), 0x00, 1); } /* DDC shift, IF */
    w.step = 16; for (uint8_t i = 0; i < 20; ++i) w.demod_write(1, (uint16_t)(0x1c + i), fir[i], 1);
    w.step = 17; w.demod_write(0, 0x19, 0x05, 1);                   /* SDR mode, DAGC off */
    w.step = 18; w.demod_write(1, 0x93, 0xf0, 1);                   /* FSM state-holding register */
    w.step = 19; w.demod_write(1, 0x94, 0x0f, 1);
    w.step = 20; w.demod_write(1, 0x11, 0x00, 1);                   /* AGC off */
    w.step = 21; w.demod_write(1, 0x04, 0x00, 1);                   /* RF and IF AGC loops off */
    w.step = 22; w.demod_write(0, 0x61, 0x60, 1);                   /* PID filter off */
    w.step = 23; w.demod_write(0, 0x06, 0x80, 1);                   /* default ADC_I / ADC_Q datapath */
    w.step = 24; w.demod_write(1, 0xb1, 0x1b, 1);                   /* zero-IF, DC cancellation, IQ compensation */
    w.step = 25; w.demod_write(0, 0x0d, 0x83, 1);                   /* no 4.096 MHz clock on TP_CK0 */
    w.step = 26; w.demod_write(1, 0x01, 0x18, 1);                   /* I2C repeater on */
    w.step = 27; w.i2c_read_reg(0xc8, 0x02);                        /* E4000 probe: E4K_I2C_ADDR, E4K_CHECK_ADDR */
    /* 28, 29: tuner Init / InitProcess -- the tuner driver's own transfers */
    w.step = 30;                                                    /* RTLSDR_set_sample_rate, states 1..9 (:700-790) */
    w.demod_write(1, 0x01, 0x18, 1);                                /* repeater on; tuner SetBW; repeater on again */
    w.demod_write(1, 0x01, 0x18, 1);
    w.demod_write(1, 0x9f, (uint16_t)(rate.rsamp_ratio >> 16), 2);
    w.demod_write(1, 0xa1, (uint16_t)(rate.rsamp_ratio & 0xffff), 2);
    w.demod_write(1, 0x3f, 0, 1);                                   /* frequency correction 0 ppm, written twice */
    w.demod_write(1, 0x3f, 0, 1);
    w.demod_write(1, 0x01, 0x14, 1);                                /* soft reset on, off */
    w.demod_write(1, 0x01, 0x10, 1);
    w.step = 31; w.demod_write(0, 0x19, (flags & B200SDR_RTL_INIT_TEST_MODE) ? 0x03 : 0x05, 1);
    w.step = 32; w.write_reg(USBB, USB_EPA_CTL, 0x1002, 2);         /* reset the bulk FIFO (mandatory) */
    w.step = 33; w.write_reg(USBB, USB_EPA_CTL, 0x0000, 2);
    *n_out = w.n;
    return w.n > capacity ? B200SDR_BUSY : B200SDR_OK;
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

/* E4K_tune_params state 4, RTL/Src/tuner_e4k.c:871-878 */
int32_t b200sdr_e4k_band(uint32_t flo_hz)
{
    if (flo_hz < 140000000u) return B200SDR_E4K_BAND_VHF2;
    if (flo_hz < 350000000u) return B200SDR_E4K_BAND_VHF3;
    if (flo_hz < 1135000000u) return B200SDR_E4K_BAND_UHF;
    return B200SDR_E4K_BAND_L;
}

namespace {
/* closest_arr_idx, tuner_e4k.c:231-247: strict '<' keeps the FIRST entry on a tie.  Tables in kHz. */
int32_t nearest_khz_entry(const uint32_t *khz, int n, uint32_t hz)
{
    int32_t best = 0;
    uint32_t best_gap = 0xFFFFFFFFu;
    for (int i = 0; i < n; ++i) {
        const uint32_t entry = khz[i] * 1000u;
        const uint32_t gap = hz > entry ? hz - entry : entry - hz;
        if (gap < best_gap) {
            best_gap = gap;
            best = i;
        }
    }
    return best;
}
const uint32_t kRfCentreUhfKhz[16] = {360000, 380000, 405000, 425000, 450000, 475000, 505000, 540000,
                                      575000, 615000, 670000, 720000, 760000, 840000, 890000, 970000};
const uint32_t kRfCentreLKhz[16] = {1300000, 1320000, 1360000, 1410000, 1445000, 1460000, 1490000, 1530000,
                                    1560000, 1590000, 1640000, 1660000, 1680000, 1700000, 1720000, 1750000};
const uint32_t kMixBwKhz[16] = {27000, 27000, 27000, 27000, 27000, 27000, 27000, 27000,
                                4600,  4200,  3800,  3400,  3300,  2700,  2300,  1900};
const uint32_t kChanBwKhz[32] = {5500, 5300, 5000, 4800, 4600, 4400, 4300, 4100, 3900, 3800, 3700,
                                 3600, 3400, 3300, 3200, 3100, 3000, 2950, 2900, 2800, 2750, 2700,
                                 2600, 2550, 2500, 2450, 2400, 2300, 2280, 2240, 2200, 2150};
const uint32_t kRcBwKhz[16] = {21400, 21000, 17600, 14700, 12400, 10600, 9000, 7700,
                               6400,  5300,  4400,  3400,  2600,  1800,  1200, 1000};
} // namespace

/* choose_rf_filter, tuner_e4k.c:250-277 over the centre tables :218-229 */
int32_t b200sdr_e4k_rf_filter(int32_t band, uint32_t freq_hz)
{
    if (band == B200SDR_E4K_BAND_UHF) return nearest_khz_entry(kRfCentreUhfKhz, 16, freq_hz);
    if (band == B200SDR_E4K_BAND_L) return nearest_khz_entry(kRfCentreLKhz, 16, freq_hz);
    return 0;
}

/* find_if_bw, tuner_e4k.c:363-372 over mix_filter_bw / ifch_filter_bw / ifrc_filter_bw :165-191 */
int32_t b200sdr_e4k_if_bw_index(int32_t filter, uint32_t bw_hz, uint32_t *actual_hz)
{
    const uint32_t *table = nullptr;
    int n = 0;
    switch (filter) {
    case B200SDR_E4K_IF_FILTER_MIX: table = kMixBwKhz; n = 16; break;
    case B200SDR_E4K_IF_FILTER_CHAN: table = kChanBwKhz; n = 32; break;
    case B200SDR_E4K_IF_FILTER_RC: table = kRcBwKhz; n = 16; break;
    default: break;
    }
    if (!table) {
        if (actual_hz) *actual_hz = 0;
        return 0;
    }
    const int32_t idx = nearest_khz_entry(table, n, bw_hz);
    if (actual_hz) *actual_hz = table[idx] * 1000u;
    return idx;
}

} /* extern "C" */
