/*
 * b200sdr_frontend.h -- RTL2832 / E4000 front-end parameter math.
 *
 * Pure host functions, no device work: what sample rate the byte stream REALLY has (exact fs for
 * the de-emphasis / resampler design), how the dongle's 32-tap decimation FIR is packed on the
 * wire, and which PLL settings (hence which exact LO frequency = the spectrum's frequency axis) an
 * E4000 gets for a requested frequency.  Each function restates one reference routine and is
 * checked bit-for-bit against the reference's own code compiled on the host (oracle A):
 *   b200sdr_rtl_resampler     RTLSDR_set_sample_rate, RTL/Src/usbh_rtlsdr.c:666-700 (ratio :683-689)
 *   b200sdr_rtl_fir_pack      RTLSDR_set_fir,         RTL/Src/usbh_rtlsdr.c:534-575
 *   b200sdr_rtl_default_fir   RTLSDR_FIR table,       RTL/Inc/usbh_rtlsdr.h:340-345
 *   b200sdr_e4k_pll_params    E4K_compute_pll_params, RTL/Src/tuner_e4k.c:689-737 (+ :301-361)
 *   b200sdr_rtl_init_sequence USBH_RTLSDR_ClassRequest, RTL/Src/usbh_rtlsdr.c:809-945 (the register writes as data)
 */
#ifndef B200SDR_FRONTEND_H
#define B200SDR_FRONTEND_H

#include <stdint.h>

#include "b200sdr.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B200SDR_RTL_XTAL_HZ 28800000u /* DEF_RTL_XTAL_FREQ, RTL/Inc/usbh_rtlsdr.h:280 */

typedef struct b200sdr_rtl_rate {
    uint32_t rsamp_ratio;      /* value written to the resampler registers                 */
    uint32_t real_rsamp_ratio; /* ratio the hardware actually applies                      */
    double   real_rate;        /* exact sample rate of the byte stream in S/s              */
} b200sdr_rtl_rate;

/* Returns B200SDR_OK and fills *out; B200SDR_NOT_SUPPORTED (out still filled) when the resampler
 * cannot produce the rate (<= 225 kS/s, > 3.2 MS/s, or in (300 k, 900 k]).  The firmware computes
 * the same numbers but loses that verdict (it overwrites FAIL with BUSY, usbh_rtlsdr.c:677-697). */
B200SDR_API int32_t b200sdr_rtl_resampler(uint32_t samp_rate, uint32_t xtal_hz, b200sdr_rtl_rate *out);

/* The 16 coefficients (outer first; 8 x int8 then 8 x int12) the firmware uploads. */
B200SDR_API void b200sdr_rtl_default_fir(int32_t coeff[16]);
/* Pack 16 coefficients into the 20 bytes written to the demod FIR registers.
 * B200SDR_NOT_SUPPORTED if a coefficient is out of range (bytes are still produced). */
B200SDR_API int32_t b200sdr_rtl_fir_pack(const int32_t coeff[16], uint8_t out20[20]);

/* One USB control transfer of the RTL2832 bring-up, exactly as the firmware puts it on the wire: the setup
 * packet (vendor request 0; RTLSDR_write_reg / RTLSDR_demod_write_reg / RTLSDR_demod_read_reg /
 * RTLSDR_write_array / RTLSDR_read_array, RTL/Src/usbh_rtlsdr.c:262-345, :445-521) and, for OUT transfers, the
 * wLength payload bytes.  `step` is the firmware's request number (case labels of usbh_rtlsdr.c:826-907). */
typedef struct b200sdr_ctl_xfer {
    uint8_t  bmRequestType; /* 0x40 vendor OUT, 0xC0 vendor IN (CTRL_OUT / CTRL_IN, RTL/Inc/usbh_rtlsdr.h:357-358) */
    uint8_t  bRequest;      /* always 0 */
    uint16_t wValue;        /* register address; demod: (addr << 8) | 0x20; I2C: the chip's bus address */
    uint16_t wIndex;        /* (block << 8) | 0x10 for writes, (block << 8) for reads; demod: page | 0x10, reads page 0x0a */
    uint16_t wLength;       /* 1 or 2 */
    uint8_t  data[2];       /* OUT payload (16-bit values travel high byte first); zero for IN */
    uint8_t  step;          /* 0..33 */
    uint8_t  reserved;
} b200sdr_ctl_xfer;

#define B200SDR_RTL_INIT_TEST_MODE 1u /* flags: step 31 switches the counter test pattern ON, as the firmware does
                                         (usbh_rtlsdr.c:901); without it the demodulator output is selected */

/* The complete control-transfer sequence USBH_RTLSDR_ClassRequest issues between enumeration and the first bulk
 * read, as data: USB block and demodulator power-up (steps 0-5), soft reset, spectrum inversion / DDC / IF
 * registers (6-15), the 20 FIR bytes (16, from `fir16`, NULL = the firmware's table), SDR mode, AGC / PID / ADC /
 * zero-IF setup (17-25), I2C repeater (26), the E4000 probe read (27), the resampler ratio for `samp_rate` with the
 * two soft resets around it (30, RTLSDR_set_sample_rate states 1-9), test mode (31) and the two EPA FIFO resets
 * (32-33).  Every demodulator write is followed by the firmware's dummy read of page 0x0a register 0x01.  The
 * tuner's own initialisation (steps 28, 29 and SetBW inside step 30) belongs to the tuner driver and is not part
 * of the list.  With samp_rate = 240000, the default FIR and B200SDR_RTL_INIT_TEST_MODE the list is transfer for
 * transfer what the reference firmware emits (108 transfers; tests/test_frontend_parity.py records the
 * reference's own FSM in oracle A).
 * Writes min(capacity, n) entries, *n_out = n.  B200SDR_OK; B200SDR_NOT_SUPPORTED for a rate the resampler cannot
 * produce or a FIR coefficient out of range (nothing is written and *n_out = 0: no list with undefined register
 * values ever leaves this function); B200SDR_BUSY if capacity < n; B200SDR_FAIL for null pointers / zero rate. */
B200SDR_API int32_t b200sdr_rtl_init_sequence(uint32_t samp_rate, uint32_t xtal_hz, const int32_t *fir16, uint32_t flags,
                                              b200sdr_ctl_xfer *out, uint32_t capacity, uint32_t *n_out);

typedef struct b200sdr_e4k_pll {
    uint32_t fosc, intended_flo, flo; /* flo = frequency actually synthesised */
    uint16_t x;
    uint8_t  z, r, r_idx, threephase;
} b200sdr_e4k_pll;

/* Returns the synthesised LO in Hz (0 if fosc is outside 16..30 MHz) and fills *out. */
B200SDR_API uint32_t b200sdr_e4k_pll_params(uint32_t fosc, uint32_t intended_flo, b200sdr_e4k_pll *out);

/* E4000 analog selection that follows the PLL in E4K_tune_params (RTL/Src/tuner_e4k.c:813-903):
 * which band, which of the 16 RF tracking filters, and which IF filter settings (hence what analog
 * passband the captured spectrum really has).  Values are enum e4k_band / enum e4k_if_filter of
 * RTL/Inc/tuner_e4k.h. */
#define B200SDR_E4K_BAND_VHF2 0
#define B200SDR_E4K_BAND_VHF3 1
#define B200SDR_E4K_BAND_UHF  2
#define B200SDR_E4K_BAND_L    3
#define B200SDR_E4K_IF_FILTER_MIX  0
#define B200SDR_E4K_IF_FILTER_CHAN 1
#define B200SDR_E4K_IF_FILTER_RC   2

/* Band for a synthesised LO: the thresholds of E4K_tune_params state 4 (tuner_e4k.c:871-878). */
B200SDR_API int32_t b200sdr_e4k_band(uint32_t flo_hz);
/* 4-bit RF filter index: choose_rf_filter (tuner_e4k.c:250-277) -- 0 on the VHF bands and for an
 * unknown band, else the first closest centre of the UHF / L table (:218-229). */
B200SDR_API int32_t b200sdr_e4k_rf_filter(int32_t band, uint32_t freq_hz);
/* Register index of the IF filter setting closest to bw_hz: find_if_bw (tuner_e4k.c:363-372) over
 * the mixer / channel / RC bandwidth tables (:165-191); 0 for an unknown filter.  *actual_hz (may be
 * NULL) receives the bandwidth that index selects (0 for an unknown filter). */
B200SDR_API int32_t b200sdr_e4k_if_bw_index(int32_t filter, uint32_t bw_hz, uint32_t *actual_hz);

#ifdef __cplusplus
}
#endif
#endif
