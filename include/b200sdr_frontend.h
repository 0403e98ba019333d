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
