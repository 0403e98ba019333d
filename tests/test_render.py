"""the 480 x 272 ARGB8888 spectrum image.  CPU: the golden definition
itself; GPU: the device kernel through the C ABI, BIT-EXACT against the golden on the same float32
power values."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_api import SYNTH_MULTITONE, Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gold_render(g, power, db_min, db_max):
    g.lib.gold_render_spectrum.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    p = np.ascontiguousarray(power, np.float32)
    img = np.zeros((272, 480), np.uint32)
    g.lib.gold_render_spectrum(p.ctypes.data, db_min, db_max, img.ctypes.data)
    return img


def test_golden_render_geometry_and_colours():
    g = Golden()
    power = np.full(1024, 1.0, np.float32)     # 0 dB everywhere
    power[0] = 1e10                            # DC: 100 dB -> centre column, full height
    power[512] = 1e5                           # -fs/2: 50 dB -> column 0, half height
    img = gold_render(g, power, 0.0, 100.0)
    assert img.shape == (272, 480) and img.dtype == np.uint32
    assert np.all(img >> 24 == 0xFF)                                   # opaque ARGB8888
    lit = (img & 0x00FFFFFF) != 0
    assert lit[:, 240].all()                                           # DC column is full height
    h0 = lit[:, 0].sum()
    assert 135 <= h0 <= 137                                            # 50 dB of 100 dB over 272 rows
    assert lit[:, 100].sum() == 1                                      # 0 dB == db_min: only T[0] <= 1.0
    assert img[271, 240] == 0xFF0000FF and img[0, 240] == 0xFFFF0300   # blue at the bottom, red at the top
    # bars are solid from the bottom
    for c in (0, 100, 240):
        col = lit[:, c]
        assert not np.any(col[:-1] & ~col[1:])


def test_render_kernels_under_host_emulation():
    """k_render_spectrum / k_render_waterfall source on the host (tests/emu): bit-exact images without a GPU."""
    g = Golden()
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "libemu_kernels.so"))  # built by conftest.py
    emu.emu_render_spectrum.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    emu.emu_render_waterfall.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(1)
    spectra = (np.abs(rng.standard_normal((3, 1024))) * 10.0 ** rng.uniform(0, 9, (3, 1024))).astype(np.float32)
    for db_min, db_max in ((0.0, 100.0), (20.0, 90.0)):
        img = np.zeros((3, 272, 480), np.uint32)
        emu.emu_render_spectrum(spectra.ctypes.data, 3, 1.0, db_min, db_max, img.ctypes.data)
        for k in range(3):
            assert np.array_equal(img[k], gold_render(g, spectra[k], db_min, db_max))
        wf = np.zeros((272, 480), np.uint32)
        emu.emu_render_waterfall(spectra.ctypes.data, 3, 1.0, db_min, db_max, wf.ctypes.data)
        assert np.array_equal(wf, gold_waterfall(g, spectra, db_min, db_max))
    # the streaming accumulator holds sums: the kernel applies 1/frames with one rounded multiply
    img = np.zeros((1, 272, 480), np.uint32)
    emu.emu_render_spectrum(spectra.ctypes.data, 1, np.float32(1.0 / 255), 0.0, 100.0, img.ctypes.data)
    assert np.array_equal(img[0], gold_render(g, spectra[0] * np.float32(1.0 / 255), 0.0, 100.0))


@pytest.mark.gpu
def test_render_bit_exact_against_golden(sdr_lib):
    g = Golden()
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 0)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        spec = s.spectrum(iq)[0]
        for db_min, db_max in ((0.0, 100.0), (20.0, 90.0), (-10.0, 60.0)):
            img = s.render_spectrum(spec, db_min, db_max)
            assert np.array_equal(img, gold_render(g, spec, db_min, db_max))
        # streaming accumulator (holds the sum; the kernel applies 1/frames like the host getter)
        s.process_samples(iq)
        spec2, frames = s.get_spectrum()
        assert frames == 255
        assert np.array_equal(s.render_spectrum(None, 0.0, 100.0), gold_render(g, spec2, 0.0, 100.0))
        rnd = np.abs(np.random.default_rng(0).standard_normal(1024)).astype(np.float32) * 1e6
        assert np.array_equal(s.render_spectrum(rnd, 10.0, 80.0), gold_render(g, rnd, 10.0, 80.0))
        with pytest.raises(sdr_lib.B200SdrError):
            s.render_spectrum(rnd, 50.0, 50.0)


# ---- waterfall (spectrogram) view ----------------------------------------------------------------
def gold_waterfall(g, spectra, db_min, db_max):
    g.lib.gold_render_waterfall.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_double, C.c_void_p]
    p = np.ascontiguousarray(spectra, np.float32).reshape(-1, 1024)
    img = np.zeros((272, 480), np.uint32)
    g.lib.gold_render_waterfall(p.ctypes.data if p.shape[0] else None, p.shape[0], db_min, db_max, img.ctypes.data)
    return img


def test_golden_waterfall_rows_and_colours():
    g = Golden()
    spectra = np.full((3, 1024), 0.5, np.float32)   # below db_min = 0 dB: black
    spectra[0, 0] = 1e10                            # row 0: DC at 100 dB -> centre column, top of the ramp
    spectra[1, 512] = 1e5                           # row 1: -fs/2 at 50 dB -> column 0, mid ramp
    spectra[2, 256] = 1.0                           # row 2: +fs/4 at exactly db_min -> bottom of the ramp
    img = gold_waterfall(g, spectra, 0.0, 100.0)
    bars = [gold_render(g, spectra[r], 0.0, 100.0) for r in range(3)]
    assert np.all(img >> 24 == 0xFF)
    assert img[0, 240] == 0xFFFF0300 and img[2, 360] == 0xFF0000FF      # red (top of ramp) / blue (bottom)
    lit = (img & 0x00FFFFFF) != 0
    assert lit.sum() == 3 and lit[0, 240] and lit[1, 0] and lit[2, 360]
    # the pixel colour is the colour of the top pixel of the bar the same power draws
    for r, c in ((0, 240), (1, 0), (2, 360)):
        col = bars[r][:, c]
        top = np.flatnonzero(col & 0x00FFFFFF)[0] if r != 2 else 271
        assert img[r, c] == col[top]
    assert np.all(img[3:] == 0xFF000000)                                 # rows >= n_rows are black
    assert np.all(gold_waterfall(g, np.zeros((0, 1024), np.float32), 0.0, 100.0) == 0xFF000000)


@pytest.mark.gpu
def test_waterfall_bit_exact_against_golden(sdr_lib):
    """A 272-row spectrogram of an FM capture: consecutive 64 KiB slices -> batch spectra -> image."""
    g = Golden()
    rows, slice_bytes = 272, 65536
    iq = g.synth(1, rows * slice_bytes, 2, 5)  # SYNTH_WBFM: the carrier visibly wanders with the message
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        spectra = s.spectrum(iq, n_captures=rows)
        assert spectra.shape == (rows, 1024)
        for db_min, db_max in ((0.0, 100.0), (30.0, 80.0)):
            assert np.array_equal(s.render_waterfall(spectra, db_min, db_max), gold_waterfall(g, spectra, db_min, db_max))
        assert np.array_equal(s.render_waterfall(spectra[:100], 0.0, 100.0), gold_waterfall(g, spectra[:100], 0.0, 100.0))
        more = np.concatenate([spectra, spectra[:30]])
        assert np.array_equal(s.render_waterfall(more, 0.0, 100.0), gold_waterfall(g, spectra, 0.0, 100.0))  # 272 shown
        assert np.all(s.render_waterfall(np.zeros((0, 1024), np.float32)) == 0xFF000000)
        img = s.render_waterfall(spectra, 30.0, 80.0)
        # the FM carrier (+50 kHz +- 75 kHz) lights columns right of centre in every row
        assert np.all(((img[:, 240:290] & 0x00FFFFFF) != 0).any(axis=1))
        with pytest.raises(sdr_lib.B200SdrError):
            s.render_waterfall(spectra, 10.0, 10.0)
