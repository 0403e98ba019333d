"""Oracle B (oracle/golden.c) checked against INDEPENDENT numpy/scipy implementations and against
the committed vectors in tests/golden/.  CPU only."""
import os

import numpy as np
import pytest
from scipy import signal

from oracle_api import (AVG_EMA, AVG_MEAN, SYNTH_AM, SYNTH_COUNTER, SYNTH_MULTITONE, SYNTH_WBFM, WIN_BLACKMAN,
                        WIN_HANN, WIN_RECT, Golden, wrap_phase)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    return Golden()


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_vectors.npz"))


def to_c(iq):
    return (iq[0::2].astype(np.float64) - 127.5) + 1j * (iq[1::2].astype(np.float64) - 127.5)


def test_convert_is_exact_half_integers(g, vec):
    out = g.convert(np.arange(256, dtype=np.uint8))
    assert np.array_equal(out, np.arange(256) - 127.5)
    assert np.array_equal(out, vec["convert_all_bytes"])
    # exactly representable in fp32 (SURVEY 8a exactness note)
    assert np.array_equal(out.astype(np.float32).astype(np.float64), out)


def test_windows(g, vec):
    n = np.arange(1024)
    assert np.allclose(g.window(WIN_HANN), 0.5 - 0.5 * np.cos(2 * np.pi * n / 1024), atol=1e-15)
    assert np.allclose(g.window(WIN_HANN), signal.get_window("hann", 1024, fftbins=True), atol=1e-15)
    assert np.allclose(g.window(WIN_BLACKMAN), signal.get_window("blackman", 1024, fftbins=True), atol=1e-15)
    assert np.array_equal(g.window(WIN_RECT), np.ones(1024))
    assert np.array_equal(g.window(WIN_HANN), vec["window_hann"])


@pytest.mark.parametrize("n", [2, 8, 32, 1024])
def test_fft_against_naive_dft_and_numpy(g, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    X = g.fft(x)
    assert np.allclose(X, np.fft.fft(x), rtol=0, atol=1e-11 * n)
    if n <= 32 or n == 1024:
        assert np.allclose(X, g.dft_naive(x), rtol=0, atol=1e-11 * n)


def test_fft_sign_convention(g):
    # a tone at +3 cycles per frame must land in bin 3 (forward transform, e^{-j...})
    n = np.arange(1024)
    X = g.fft(np.exp(2j * np.pi * 3 * n / 1024))
    assert np.argmax(np.abs(X)) == 3


@pytest.mark.parametrize("window,name", [(WIN_HANN, "hann"), (WIN_BLACKMAN, "blackman"), (WIN_RECT, "boxcar")])
def test_spectrum_mean_against_numpy(g, window, name):
    iq = g.synth(1, 2 * (1024 + 512 * 20 + 100), SYNTH_MULTITONE, 5)
    P, frames = g.spectrum(iq, window=window)
    x = to_c(iq)
    assert frames == (x.size - 1024) // 512 + 1 == 21
    w = signal.get_window(name, 1024, fftbins=True)
    ref = np.mean([np.abs(np.fft.fft(x[512 * m:512 * m + 1024] * w)) ** 2 for m in range(frames)], axis=0)
    assert np.allclose(P, ref, rtol=1e-11)


def test_spectrum_ema_against_numpy(g):
    iq = g.synth(1, 2 * (1024 + 512 * 9), SYNTH_MULTITONE, 6)
    P, frames = g.spectrum(iq, avg_mode=AVG_EMA, beta=0.1)
    x, w, acc = to_c(iq), signal.get_window("hann", 1024, fftbins=True), np.zeros(1024)
    for m in range(frames):
        acc = 0.9 * acc + 0.1 * np.abs(np.fft.fft(x[512 * m:512 * m + 1024] * w)) ** 2
    assert np.allclose(P, acc, rtol=1e-11)


def test_spectrum_short_input_is_empty(g):
    P, frames = g.spectrum(np.zeros(2 * 1023, np.uint8))
    assert frames == 0 and not P.any()


def test_multitone_lands_on_its_bins(g):
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 0)  # BASELINE config[0] block
    P, frames = g.spectrum(iq)
    assert frames == 255
    top = set(np.argsort(P)[-12:].tolist())
    # -600 kHz, -150 kHz, +37.5 kHz, +450 kHz at fs = 2.4 MHz -> bins 768, 960, 16, 192 (+ Hann neighbours)
    assert {768, 960, 16, 192} <= top


SCIPY_TAPS = [
    (80, 100e3, 2.4e6, 8.0, 1 / 127.5),
    (50, 16e3, 240e3, 5.0, 240000.0 / (2 * np.pi * 75000.0)),
    (80, 55e3, 2.4e6, 6.0, 1 / 127.5),
    (120, 5.0e3, 120e3, 6.0, 1.0),
    (48, 3.6e3, 24e3, 6.0, 2.0),
]


@pytest.mark.parametrize("which", range(5))
def test_taps_against_scipy_firwin(g, vec, which):
    n, fc, fs, beta, gain = SCIPY_TAPS[which]
    h = g.taps(which)
    ref = signal.firwin(n, fc, window=("kaiser", beta), fs=fs, scale=False)
    ref = ref / ref.sum() * gain
    assert h.size == n
    assert np.allclose(h, ref, rtol=0, atol=1e-15 * abs(gain) * 10 + 1e-17)
    assert np.allclose(h, h[::-1], rtol=0, atol=1e-18)  # linear phase: the kernels rely on it
    assert np.array_equal(h, vec[f"taps{which}"])


def scipy_wbfm(iq, g):
    x = to_c(iq)
    h1, h2 = g.taps(0), g.taps(1)
    y1 = signal.lfilter(h1, 1.0, x)[::10]
    d = np.angle(y1 * np.conj(np.concatenate([[0], y1[:-1]])))
    d[0] = 0.0
    alpha = 1 - np.exp(-1 / 18)
    e = signal.lfilter([alpha], [1, -(1 - alpha)], d)
    return signal.lfilter(h2, 1.0, e)[::5], d


def test_wbfm_against_scipy(g):
    iq = g.synth(1, 2 * 30011, SYNTH_WBFM, 9)
    audio, disc = g.wbfm(iq, want_disc=True)
    ref_a, ref_d = scipy_wbfm(iq, g)
    assert audio.size == ref_a.size and disc.size == ref_d.size
    assert np.max(np.abs(wrap_phase(disc - ref_d))) < 1e-12
    assert np.max(np.abs(audio - ref_a)) < 1e-12


def test_wbfm_recovers_the_message(g):
    # 1 kHz at 0.5 and 5 kHz at 0.3 of full deviation, plus the +50 kHz carrier offset as DC (50/75)
    iq = g.synth(1, 2 * 240000, SYNTH_WBFM, 1)
    a = g.wbfm(iq)[2000:]
    t = np.arange(a.size) / 48000.0
    A = np.stack([np.ones_like(t), np.sin(2 * np.pi * 1e3 * t), np.cos(2 * np.pi * 1e3 * t),
                  np.sin(2 * np.pi * 5e3 * t), np.cos(2 * np.pi * 5e3 * t)], axis=1)
    c, *_ = np.linalg.lstsq(A, a, rcond=None)
    assert abs(c[0] - 50 / 75) < 0.01
    assert abs(np.hypot(c[1], c[2]) - 0.5 * abs(1 / (1 + 1j * 2 * np.pi * 1e3 * 75e-6))) < 0.02
    assert np.hypot(c[3], c[4]) > 0.05


def scipy_am(iq, g):
    x = to_c(iq)
    g1, g2, g3 = g.taps(2), g.taps(3), g.taps(4)
    y1 = signal.lfilter(g1, 1.0, x)[::20]
    y2 = signal.lfilter(g2, 1.0, y1)[::10]
    r = np.abs(y2)
    b = signal.lfilter([1, -1], [1, -0.999], r)
    v = np.zeros(2 * b.size)
    v[::2] = b
    return signal.lfilter(g3, 1.0, v)[::3]


def test_am_against_scipy(g):
    iq = g.synth(1, 2 * 200017, SYNTH_AM, 4)
    audio = g.am(iq)
    ref = scipy_am(iq, g)
    assert audio.size == ref.size
    assert np.max(np.abs(audio - ref)) < 1e-13


def test_float32_build_tracks_float64(g):
    g32 = Golden(f32=True)
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 0)
    P64, _ = g.spectrum(iq)
    P32, _ = g32.spectrum(iq)
    assert np.max(np.abs(P32 - P64) / P64) < 2e-4
    iq = g.synth(1, 2 * 24000, SYNTH_WBFM, 0)
    assert np.max(np.abs(g32.wbfm(iq) - g.wbfm(iq))) < 1e-4


def test_committed_vectors_reproduce(g, vec):
    iq = g.synth(1, int(vec["spec_len"]), SYNTH_MULTITONE, int(vec["spec_seed"]))
    assert np.array_equal(g.spectrum(iq)[0], vec["spec_hann_mean"])
    assert np.array_equal(g.spectrum(iq, window=WIN_BLACKMAN)[0], vec["spec_blackman_mean"])
    assert np.array_equal(g.spectrum(iq, avg_mode=AVG_EMA, beta=0.1)[0], vec["spec_hann_ema"])
    iq = g.synth(1, int(vec["fm_len"]), SYNTH_WBFM, int(vec["fm_seed"]))
    a, d = g.wbfm(iq, want_disc=True)
    assert np.array_equal(a, vec["fm_audio"]) and np.array_equal(d, vec["fm_disc"])
    iq = g.synth(1, int(vec["am_len"]), SYNTH_AM, int(vec["am_seed"]))
    assert np.array_equal(g.am(iq), vec["am_audio"])
    for kind, name in ((SYNTH_COUNTER, "counter"), (SYNTH_MULTITONE, "multitone"), (SYNTH_WBFM, "wbfm"), (SYNTH_AM, "am")):
        assert np.array_equal(g.synth(1, 256, kind, 7), vec[f"synth_{name}_head"])


def test_counter_stream_is_the_rtl2832_test_pattern(g):
    # test mode (reference usbh_rtlsdr.c:901): an 8-bit counter; b[i+1] == b[i] + 1 mod 256
    b = g.synth(1, 4096, SYNTH_COUNTER, 0).astype(np.int32)
    assert np.all((b[1:] - b[:-1]) % 256 == 1) and b[0] == 0


def test_counter_check_against_numpy(g):
    """gold_counter_check (the checker of kernel K0) against a two-line numpy statement of the same definition."""
    rng = np.random.default_rng(5)
    for n in (0, 1, 4, 255, 256, 257, 4096, 100003):
        u = (np.arange(n) + 37).astype(np.uint8)
        hits = rng.integers(0, max(n, 1), size=min(n, 7))
        u[hits] = rng.integers(0, 256, size=hits.size, dtype=np.uint8) if n else u[hits]
        for expect in (-1, 37, 38):
            bad = np.zeros(n, bool)
            if n:
                bad[1:] = u[1:] != (u[:-1].astype(np.int64) + 1) % 256
                bad[0] = expect >= 0 and u[0] != expect
            idx = np.flatnonzero(bad)
            assert g.counter_check(u, expect) == (idx.size, int(idx[0]) if idx.size else None)
