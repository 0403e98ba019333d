/*
 * filter_design.h -- host-side design of the fixed filters, window and scan constants the
 * device kernels use (double precision, rounded to float once).  Product code; the oracle
 * (oracle/golden.c) restates the same formulas independently and tests compare the two through
 * b200sdr_get_taps().
 *
 * Parameters (frozen):
 *   FM1  80 taps  fc 100 kHz @ 2.4 MS/s   Kaiser beta 8  gain 1/127.5
 *   FM2  50 taps  fc 16 kHz  @ 240 kS/s   Kaiser beta 5  gain 240000 / (2 pi 75000)
 *   AM1  80 taps  fc 55 kHz  @ 2.4 MS/s   Kaiser beta 6  gain 1/127.5
 *   AM2  120 taps fc 5.0 kHz @ 120 kS/s   Kaiser beta 6  gain 1
 *   AM3  48 taps  fc 3.6 kHz @ 24 kS/s    Kaiser beta 6  gain 2   (x2 interpolation)
 * The RTL2832's own decimation FIR, whose coefficients the firmware uploads
 * (RTL/Inc/usbh_rtlsdr.h:340-345, RTL/Src/usbh_rtlsdr.c:534-613), sits in front of all of these.
 */
#ifndef B200_FILTER_DESIGN_H
#define B200_FILTER_DESIGN_H

#include <cmath>
#include <vector>

namespace b200 {

constexpr double kPi = 3.14159265358979323846;

inline double bessel_i0(double x)
{
    /* power series sum ((x/2)^k / k!)^2 */
    double q = 0.25 * x * x, term = 1.0, sum = 1.0;
    for (int k = 1; k < 400; ++k) {
        term *= q / (double(k) * double(k));
        sum += term;
        if (term < sum * 1e-21) break;
    }
    return sum;
}

/* Kaiser-windowed sinc low-pass, normalised so the taps sum to `gain` */
inline std::vector<double> kaiser_lowpass(int ntaps, double fc_cycles_per_sample, double beta, double gain)
{
    std::vector<double> h(ntaps);
    const double centre = 0.5 * (ntaps - 1), norm = bessel_i0(beta);
    double total = 0.0;
    for (int k = 0; k < ntaps; ++k) {
        const double t = k - centre, x = 2.0 * fc_cycles_per_sample * t;
        const double ideal = std::fabs(x) < 1e-12 ? 1.0 : std::sin(kPi * x) / (kPi * x);
        const double r = t / centre;
        const double inside = r * r < 1.0 ? 1.0 - r * r : 0.0;
        h[k] = 2.0 * fc_cycles_per_sample * ideal * bessel_i0(beta * std::sqrt(inside)) / norm;
        total += h[k];
    }
    for (double &v : h) v *= gain / total;
    return h;
}

inline std::vector<double> design_taps(unsigned which)
{
    switch (which) {
    case 0: return kaiser_lowpass(80, 100000.0 / 2400000.0, 8.0, 1.0 / 127.5);
    case 1: return kaiser_lowpass(50, 16000.0 / 240000.0, 5.0, 240000.0 / (2.0 * kPi * 75000.0));
    case 2: return kaiser_lowpass(80, 55000.0 / 2400000.0, 6.0, 1.0 / 127.5);
    case 3: return kaiser_lowpass(120, 5000.0 / 120000.0, 6.0, 1.0);
    case 4: return kaiser_lowpass(48, 3600.0 / 24000.0, 6.0, 2.0);
    default: return {};
    }
}

inline double deemph_alpha() { return 1.0 - std::exp(-1.0 / (240000.0 * 75e-6)); }
inline double dcblock_rho() { return 0.999; }

/* periodic windows of length n: 0 rectangular, 1 Hann, 2 Blackman */
inline std::vector<float> make_window(unsigned kind, int n)
{
    std::vector<float> w(n);
    for (int i = 0; i < n; ++i) {
        const double a = 2.0 * kPi * i / n;
        double v = 1.0;
        if (kind == 1) v = 0.5 - 0.5 * std::cos(a);
        else if (kind == 2) v = 0.42 - 0.5 * std::cos(a) + 0.08 * std::cos(2.0 * a);
        w[i] = (float)v;
    }
    return w;
}

} // namespace b200
#endif
