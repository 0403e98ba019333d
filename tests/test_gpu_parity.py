"""Parity tests proper: the CUDA path, called through the C ABI (libb200sdr.so), against the
oracles on the same seeded inputs.  Run on the B200 box: `pytest -m gpu`.

Tolerances (BASELINE.json north_star):
  * u8 -> cf32 conversion, ingest bytes, block indexing: BIT-EXACT
  * spectrum power: <= 1e-5 relative error per bin
  * discriminator: <= 1e-4 rad (compared modulo 2 pi); audio: <= 1e-4 x gain
"""
import os

import numpy as np
import pytest

from oracle_api import (AVG_EMA, SYNTH_AM, SYNTH_COUNTER, SYNTH_MULTITONE, SYNTH_WBFM, WIN_BLACKMAN, WIN_HANN,
                        WIN_RECT, Golden, RefHost, wrap_phase, extreme_patterns)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SPEC_RTOL = 1e-5      # relative, per bin
DISC_ATOL = 1e-4      # radians
FM_AUDIO_ATOL = 1e-4 * 240000.0 / (2 * np.pi * 75000.0)
AM_AUDIO_ATOL = 1e-6  # AM audio is O(0.05); fp32 FIRs give ~1e-8


@pytest.fixture(scope="module")
def g():
    return Golden()


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_vectors.npz"))


@pytest.fixture(scope="module")
def sdr(sdr_lib):
    s = sdr_lib.B200Sdr()
    yield s
    s.close()


# ------------------------------------------------------------------------------------------ K2
def test_convert_bit_exact_all_bytes(sdr, g, vec):
    iq = np.arange(256, dtype=np.uint8)
    out = sdr.convert_cf32(iq)
    assert out.dtype == np.float32
    assert np.array_equal(out.astype(np.float64), vec["convert_all_bytes"])  # exact: u - 127.5


@pytest.mark.parametrize("nbytes", [4, 12, 16, 1020, 262144])
def test_convert_bit_exact_ragged(sdr, g, nbytes):
    iq = np.random.default_rng(nbytes).integers(0, 256, nbytes, dtype=np.uint8)
    assert np.array_equal(sdr.convert_cf32(iq).astype(np.float64), g.convert(iq))


@pytest.mark.parametrize("window", [WIN_HANN, WIN_BLACKMAN])
def test_convert_windowed(sdr, sdr_lib, g, window):
    iq = g.synth(1, 8192, SYNTH_MULTITONE, 1)
    out = sdr.convert_cf32(iq, window=window)
    # device window is the float32 rounding of the golden window; product is one fp32 multiply
    w32 = g.window(window).astype(np.float32)
    assert np.array_equal(sdr.get_window(window), w32)
    expect = (g.convert(iq).astype(np.float32).reshape(-1, 2) * np.tile(w32, 4)[: iq.size // 2, None]).reshape(-1)
    assert np.array_equal(out, expect)  # bit-exact against the same fp32 multiply
    assert np.max(np.abs(out - g.convert(iq, window=window))) <= 128 * 2.0 ** -23


def test_taps_match_oracle(sdr, g):
    for which in range(5):
        assert np.array_equal(sdr.get_taps(which), g.taps(which).astype(np.float32))


# ------------------------------------------------------------------------------------ spectrum
def spec_check(out, gold, rtol=SPEC_RTOL):
    rel = np.abs(out.astype(np.float64) - gold) / gold
    assert rel.max() <= rtol, f"max rel err {rel.max():.3e} at bin {rel.argmax()}"


def spec_check_few_frames(out, gold):
    """With only a handful of frames nothing averages the fp32 rounding noise of the strongest
    component down, so a weak bin next to a strong tone cannot meet 1e-5 of ITS OWN power: the
    achievable bound is eps * sqrt(P_k * P_max).  (The BASELINE configs average 255 / 46 874
    frames and are held to the plain 1e-5.)"""
    err = np.abs(out.astype(np.float64) - gold)
    bound = 1e-5 * gold + 4e-7 * np.sqrt(gold * gold.max())
    assert np.all(err <= bound), f"worst excess {np.max(err / bound):.2f}x at bin {np.argmax(err / bound)}"


def test_spectrum_config0_block(sdr, g):
    """BASELINE config[0]: one 256 KiB block -> 1024-pt Hann FFT power spectrum (255 frames)."""
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 0)
    out = sdr.spectrum(iq)
    gold, frames = g.spectrum(iq)
    assert frames == 255
    spec_check(out[0], gold)


def test_spectrum_golden_fixture(sdr, sdr_lib, vec):
    iq = sdr_lib.synth_fill_host(1, int(vec["spec_len"]), SYNTH_MULTITONE, int(vec["spec_seed"]))
    spec_check_few_frames(sdr.spectrum(iq)[0], vec["spec_hann_mean"])  # only 15 frames averaged


@pytest.mark.parametrize("n_captures,len_each", [(1, 2048), (3, 2048 + 1024 * 7 + 16), (5, 262144), (2, 4 * 262144)])
def test_spectrum_batches(sdr, g, n_captures, len_each):
    iq = g.synth(n_captures, len_each, SYNTH_MULTITONE, 40)
    out = sdr.spectrum(iq, n_captures)
    for c in range(n_captures):
        gold, frames = g.spectrum(iq[c * len_each:(c + 1) * len_each])
        spec_check(out[c], gold) if frames >= 200 else spec_check_few_frames(out[c], gold)



def test_spectrum_cufft_cross_check(sdr, g):
    """north_star: "cuFFT is used only as a cross-check, never on the product path".  A full 10 s
    capture through torch.fft (cuFFT Z2Z, float64), against the product spectrum AND the golden
    model -- three independent FFT implementations agreeing."""
    torch = pytest.importorskip("torch")
    len_each = 48_000_000
    iq = g.synth(1, len_each, SYNTH_MULTITONE, 77)
    out = sdr.spectrum(iq)[0].astype(np.float64)
    w = torch.from_numpy(0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(1024) / 1024.0)).cuda()
    assert np.max(np.abs(sdr.get_window(WIN_HANN) - w.cpu().numpy())) <= 6e-8  # the table the kernel uses
    x = torch.from_numpy(iq.reshape(-1, 2)).cuda().to(torch.float64) - 127.5
    frames = torch.complex(x[:, 0], x[:, 1]).unfold(0, 1024, 512)
    assert frames.shape[0] == 46874
    acc = torch.zeros(1024, dtype=torch.float64, device="cuda")
    for lo in range(0, frames.shape[0], 8192):  # bounded temporaries
        acc += (torch.fft.fft(frames[lo:lo + 8192] * w, dim=1).abs() ** 2).sum(0)
    cufft = (acc / frames.shape[0]).cpu().numpy()
    spec_check(out, cufft)
    gold, _ = g.spectrum(iq)
    assert np.max(np.abs(gold - cufft) / cufft) <= 1e-9  # float64 golden vs float64 cuFFT


def test_spectrum_short_capture_is_zero(sdr):
    assert not sdr.spectrum(np.zeros(2032, np.uint8)).any()  # < 1024 samples: no frame


def test_spectrum_random_bytes(sdr, g):
    iq = np.random.default_rng(0).integers(0, 256, 262144, dtype=np.uint8)
    spec_check(sdr.spectrum(iq)[0], g.spectrum(iq)[0])


def test_spectrum_extreme_bytes(sdr, g):
    iq = np.zeros(65536, np.uint8)
    iq[::3] = 255  # full-scale square-ish pattern: largest magnitudes the FFT can see
    out, gold = sdr.spectrum(iq)[0].astype(np.float64), g.spectrum(iq)[0]
    assert np.max(np.abs(out - gold)) <= 1e-5 * gold.max()


@pytest.mark.parametrize("window", [WIN_RECT, WIN_BLACKMAN])
def test_spectrum_other_windows(sdr_lib, g, window):
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 3)
    with sdr_lib.B200Sdr(window=window) as s:
        out = s.spectrum(iq)[0].astype(np.float64)
    gold = g.spectrum(iq, window=window)[0]
    assert np.max(np.abs(out - gold) / np.maximum(gold, 1e-7 * gold.max())) <= 1e-5


def test_spectrum_ema(sdr_lib, g):
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 4)
    with sdr_lib.B200Sdr(avg_mode=AVG_EMA, ema_beta=0.1) as s:
        out = s.spectrum(iq)[0]
    spec_check_few_frames(out, g.spectrum(iq, avg_mode=AVG_EMA, beta=0.1)[0])  # ~20 effective frames


def test_spectrum_parseval(sdr, g):
    """Size-independent property: sum_k P[k] = 1024 * mean_m sum_n |w[n] x_m[n]|^2."""
    iq = g.synth(1, 2 * 262144, SYNTH_MULTITONE, 8)
    P = sdr.spectrum(iq)[0].astype(np.float64)
    x = g.convert(iq).reshape(-1, 2)
    x = x[:, 0] + 1j * x[:, 1]
    w = g.window(WIN_HANN)
    frames = (x.size - 1024) // 512 + 1
    e = np.mean([np.sum(np.abs(x[512 * m:512 * m + 1024] * w) ** 2) for m in range(frames)])
    assert abs(P.sum() - 1024 * e) <= 1e-6 * 1024 * e


# ---------------------------------------------------------------------------------------- WBFM
def test_wbfm_golden_fixture(sdr, sdr_lib, vec):
    iq = sdr_lib.synth_fill_host(1, int(vec["fm_len"]), SYNTH_WBFM, int(vec["fm_seed"]))
    audio, disc = sdr.wbfm(iq, want_disc=True)
    assert np.max(np.abs(wrap_phase(disc[0] - vec["fm_disc"]))) <= DISC_ATOL
    assert np.max(np.abs(audio[0] - vec["fm_audio"])) <= FM_AUDIO_ATOL


@pytest.mark.parametrize("n_captures,len_each", [(1, 16), (1, 2416), (2, 262144), (3, 30720 * 5 + 240 * 3 + 16), (1, 4800000)])
def test_wbfm_batches(sdr, g, n_captures, len_each):
    iq = g.synth(n_captures, len_each, SYNTH_WBFM, 50)
    audio, disc = sdr.wbfm(iq, n_captures, want_disc=True)
    for c in range(n_captures):
        ga, gd = g.wbfm(iq[c * len_each:(c + 1) * len_each], want_disc=True)
        assert audio[c].size == ga.size and disc[c].size == gd.size
        assert np.max(np.abs(wrap_phase(disc[c] - gd))) <= DISC_ATOL
        assert np.max(np.abs(audio[c] - ga)) <= FM_AUDIO_ATOL


@pytest.mark.parametrize("name", ["all0", "all255", "half_lsb", "split", "square", "fullscale_tone", "half_lsb_rotation", "tone_on_dc"])
def test_fir_chains_extreme_bytes(sdr, g, name):
    """The FIR kernels read the bytes as raw (sub)normal floats and cancel the -127.5 offset in the accumulators
    (cplx2.cuh form C): largest offsets, smallest possible signals, both halves of the subnormal / normal split."""
    nb = 2 * (30720 * 3 + 1208)
    const, rot = extreme_patterns(nb)
    iq = rot[name] if name in rot else const[name]
    if name in rot:  # a constant input sits on the discriminator's branch cut: FM only for the rotating patterns
        audio, disc = sdr.wbfm(iq, 1, want_disc=True)
        ga, gd = g.wbfm(iq, want_disc=True)
        # the first outputs see the filter's rise from x[n < 0] = 0, where |y| passes through ~0
        assert np.max(np.abs(wrap_phase(disc[0] - gd))[16:]) <= DISC_ATOL
        assert np.max(np.abs(audio[0] - ga)[8:]) <= FM_AUDIO_ATOL
    assert np.max(np.abs(sdr.am(iq, 1)[0] - g.am(iq))) <= 2 * AM_AUDIO_ATOL


def test_wbfm_host_path_equals_device_path(sdr, g):
    iq = g.synth(4, 262144, SYNTH_WBFM, 60)
    a_host = sdr.wbfm(iq, 4)
    a_dev, _ = sdr.wbfm(iq, 4, want_disc=True)
    assert np.array_equal(a_host, a_dev)


# ------------------------------------------------------------- WBFM, tensor-core FIR engine (K4-FM-TC)
@pytest.fixture(scope="module")
def sdr_tc(sdr_lib):
    s = sdr_lib.B200Sdr(fir_engine=sdr_lib.FIR_ENGINE_TENSOR)
    yield s
    s.close()


def test_tensor_engine_accumulators_bit_exact(sdr_tc, g):
    """csrc/wbfm_tc.cuh: the tcgen05.mma kind::i8 product of the raw bytes (TMA-loaded, 64B swizzle) with the three signed
    8-bit tap slices is an integer FIR; its first tile must equal numpy's integer convolution bit for bit (this pins the
    tensor map, both shared-memory operand layouts, the descriptors and the accumulator addressing)."""
    n = 160 * 125 * 2 + 1232
    iq = g.synth(1, 2 * n, SYNTH_WBFM, 11)
    acc, q, e = sdr_tc.debug_wbfm_tc_acc(iq)
    h = g.taps(0)
    hq = (q[0] * 2.0**-7 + q[1] * 2.0**-14 + q[2] * 2.0**-21) / 2.0**e
    assert np.max(np.abs(hq - h)) < 4e-7 * np.max(np.abs(h))                # the slices ARE the oracle's taps to 21 bits
    u = iq.astype(np.int64)
    for s_ in range(3):
        for c in (0, 1):
            x = np.concatenate([np.zeros(96, np.int64), u[c::2]])           # x[n < 0] = 0
            full = np.convolve(x, q[s_].astype(np.int64))                    # full[96 + n] = sum_t q[t] x[n - t]
            for j in range(17):                                              # output m = 16 r - 1 + j, sample 10 m
                m = 16 * np.arange(125) - 1 + j
                assert np.array_equal(acc[:125, 34 * s_ + 2 * j + c], full[96 + 10 * m]), (s_, c, j)


@pytest.mark.parametrize("n_captures,len_each", [(1, 16), (1, 2416), (2, 262144), (3, 30720 * 5 + 240 * 3 + 16), (1, 4800000),
                                                 (2, 320 * 125 * 3), (5, 320 * 130 + 16), (1, 320), (300, 4096), (40, 3200)])
def test_tensor_engine_batches(sdr_tc, g, n_captures, len_each):
    """same cases as test_wbfm_batches plus lengths that are / are not whole 320-byte rows (TMA tiles vs the
    bounds-checked cp.async fill of a capture's ragged last tile) and several captures per CTA"""
    iq = g.synth(n_captures, len_each, SYNTH_WBFM, 50)
    audio, disc = sdr_tc.wbfm(iq, n_captures, want_disc=True)
    for c in range(n_captures):
        ga, gd = g.wbfm(iq[c * len_each:(c + 1) * len_each], want_disc=True)
        assert audio[c].size == ga.size and disc[c].size == gd.size
        assert np.max(np.abs(wrap_phase(disc[c] - gd))) <= DISC_ATOL
        assert np.max(np.abs(audio[c] - ga)) <= FM_AUDIO_ATOL


def test_tensor_engine_golden_fixture_and_engines_agree(sdr, sdr_tc, sdr_lib, vec, g):
    iq = sdr_lib.synth_fill_host(1, int(vec["fm_len"]), SYNTH_WBFM, int(vec["fm_seed"]))
    audio, disc = sdr_tc.wbfm(iq, want_disc=True)
    assert np.max(np.abs(wrap_phase(disc[0] - vec["fm_disc"]))) <= DISC_ATOL
    assert np.max(np.abs(audio[0] - vec["fm_audio"])) <= FM_AUDIO_ATOL
    a32, d32 = sdr.wbfm(iq, want_disc=True)
    assert 0 < np.max(np.abs(audio - a32)) <= 2e-6 and np.max(np.abs(wrap_phase(disc[0] - d32[0]))[8:]) <= 5e-6
    # the host-buffer path (b200sdr_batch_host) runs the same engine
    iq4 = g.synth(4, 262144, SYNTH_WBFM, 60)
    assert np.array_equal(sdr_tc.wbfm(iq4, 4), sdr_tc.wbfm(iq4, 4, want_disc=True)[0])


@pytest.mark.parametrize("name", ["fullscale_tone", "half_lsb_rotation", "tone_on_dc"])
def test_tensor_engine_extreme_bytes(sdr_tc, g, name):
    """the offset is removed in exact integer / half-integer arithmetic: even the smallest phasor a u8 stream can
    carry stays far inside the bound (3e-7 rad measured, 6.5e-5 for the FP32 engine)"""
    nb = 2 * (30720 * 3 + 1208)
    iq = extreme_patterns(nb)[1][name]
    audio, disc = sdr_tc.wbfm(iq, 1, want_disc=True)
    ga, gd = g.wbfm(iq, want_disc=True)
    assert np.max(np.abs(wrap_phase(disc[0] - gd))[16:]) <= 2e-5
    assert np.max(np.abs(audio[0] - ga)[8:]) <= 2e-5


def test_tensor_engine_full_size_capture_alone_equals_in_batch(sdr_tc, g):
    """a 10 s capture (1200 tiles, many segments with pre-roll) against the oracle, and bitwise the same audio alone
    and inside a batch that changes the segment plan"""
    len_each = 48_000_000
    iq = g.synth(1, len_each, SYNTH_WBFM, 1001)
    alone, disc = sdr_tc.wbfm(iq, 1, want_disc=True)
    ga, gd = g.wbfm(iq, want_disc=True)
    assert np.max(np.abs(alone[0] - ga)) <= FM_AUDIO_ATOL and np.max(np.abs(wrap_phase(disc[0] - gd))) <= DISC_ATOL
    both = sdr_tc.wbfm(np.concatenate([g.synth(1, len_each, SYNTH_WBFM, 7), iq]), 2)
    # segments after the first pre-roll one tile (de-emphasis pole 0.946^2000 = 0): the plan changes where they start
    assert np.max(np.abs(both[1] - alone[0])) <= 1e-6


def test_fir_engine_config_is_validated(sdr_lib):
    with pytest.raises(sdr_lib.B200SdrError) as ei:
        sdr_lib.B200Sdr(fir_engine=2)
    assert ei.value.status == sdr_lib.NOT_SUPPORTED


# ------------------------------------------------------------------------- test-mode counter (K0)
@pytest.mark.parametrize("nbytes", [0, 4, 12, 16, 20, 1020, 16384 + 8, 262144, 4 * 262144 + 4])
def test_counter_check_host_blocks(sdr, g, nbytes):
    """bit-exact against gold_counter_check: clean counter, injected breaks (also on word, vector, warp and block
    boundaries and on the last byte), first-byte expectation."""
    rng = np.random.default_rng(nbytes)
    clean = (np.arange(nbytes) + 201).astype(np.uint8)
    assert sdr.counter_check(clean) == (0, None) == g.counter_check(clean)
    assert sdr.counter_check(clean, 201 if nbytes else -1) == (0, None)
    if nbytes == 0:
        return
    assert sdr.counter_check(clean, 7) == (1, 0) == g.counter_check(clean, 7)
    for _ in range(4):
        u = clean.copy()
        where = list(rng.integers(0, nbytes, size=5)) + [nbytes - 1, 3, 4, 15, 16, 511, 512, 16383, 16384]
        where = [w for w in where if w < nbytes]
        u[where] = rng.integers(0, 256, size=len(where), dtype=np.uint8)
        for expect in (-1, int(clean[0])):
            assert sdr.counter_check(u, expect) == g.counter_check(u, expect)


def test_counter_check_batch_dev(sdr, g):
    """the device generator's counter captures are clean; every capture is judged on its own (no carry from the
    previous capture's last byte); a dropped sector shows up as one break at the right place."""
    n_captures, len_each = 5, 3 * 262144 + 16
    d = sdr.dev_alloc(n_captures * len_each)
    try:
        sdr.synth_fill_dev(d, n_captures, len_each, SYNTH_COUNTER, 0)
        n, f = sdr.counter_check_dev(d, n_captures, len_each)
        assert not n.any() and (f == np.uint64(2**64 - 1)).all()
        host = sdr.to_host(d, n_captures * len_each, np.uint8).reshape(n_captures, len_each).copy()
        host[2, 1000:] = np.roll(host[2], -32)[1000:]  # 32 bytes lost at index 1000 (and the rolled-in head near the end)
        host[4, 70000] ^= 0x40
        sdr.to_dev(d, host.reshape(-1))
        n, f = sdr.counter_check_dev(d, n_captures, len_each)
        for c in range(n_captures):
            want_n, want_f = g.counter_check(host[c])
            assert (int(n[c]), None if f[c] == np.uint64(2**64 - 1) else int(f[c])) == (want_n, want_f), c
        assert int(n[2]) >= 1 and int(f[2]) == 1000 and int(n[4]) == 2 and int(f[4]) == 70000
    finally:
        sdr.dev_free(d)


def test_counter_check_rejects_bad_arguments(sdr, sdr_lib):
    d = sdr.dev_alloc(64)
    try:
        for args in ((d, 1, 6), (d + 4, 1, 16), (d, 2, 20)):
            with pytest.raises(sdr_lib.B200SdrError):
                sdr.counter_check_dev(*args)
    finally:
        sdr.dev_free(d)


# ------------------------------------------------------------------------------------------ AM
def test_am_golden_fixture(sdr, sdr_lib, vec):
    iq = sdr_lib.synth_fill_host(1, int(vec["am_len"]), SYNTH_AM, int(vec["am_seed"]))
    assert np.max(np.abs(sdr.am(iq)[0] - vec["am_audio"])) <= AM_AUDIO_ATOL


@pytest.mark.parametrize("n_captures,len_each", [(1, 16), (1, 4816), (2, 262144), (2, 51200 * 3 + 400 * 7 + 32), (1, 4800000)])
def test_am_batches(sdr, g, n_captures, len_each):
    iq = g.synth(n_captures, len_each, SYNTH_AM, 70)
    audio = sdr.am(iq, n_captures)
    for c in range(n_captures):
        ga = g.am(iq[c * len_each:(c + 1) * len_each])
        assert audio[c].size == ga.size
        assert np.max(np.abs(audio[c] - ga)) <= AM_AUDIO_ATOL


# ----------------------------------------------------------------------------------- streaming
def feed(s, iq, cuts):
    pos = 0
    for n in cuts:
        rc = s.process_samples(iq[pos:pos + n], allow_busy=True)
        while rc == 1:  # B200SDR_BUSY: ring full -> call again (cooperative, like USBH_BUSY)
            s.sync()
            rc = s.process_samples(iq[pos:pos + n], allow_busy=True)
        pos += n
    assert pos == iq.size


def random_cuts(total, max_block, seed):
    rng = np.random.default_rng(seed)
    cuts, left = [], total
    while left:
        n = int(min(left, 4 * rng.integers(1, max_block // 4 + 1)))
        cuts.append(n)
        left -= n
    return cuts


@pytest.mark.parametrize("submit_bytes", [0, 4], ids=["coalesced", "per_block"])
@pytest.mark.parametrize("cuts_kind", ["baseline_blocks", "reference_512", "random"])
def test_streaming_all_chains_block_cut_invariance(sdr_lib, g, cuts_kind, submit_bytes):
    """A stream cut into blocks at arbitrary 4-byte boundaries gives the single-capture result,
    whether blocks are coalesced into full ring slots (default) or submitted one by one."""
    total = 262144 * 3 + 512 * 5
    iq = g.synth(1, total, SYNTH_WBFM, 90)
    cuts = {"baseline_blocks": [262144] * 3 + [512 * 5], "reference_512": [512] * (total // 512),
            "random": random_cuts(total, 65536, 7)}[cuts_kind]
    with sdr_lib.B200Sdr(slot_bytes=262144, ring_slots=4, submit_bytes=submit_bytes) as s:
        feed(s, iq, cuts)
        spec, frames = s.get_spectrum()
        fm = s.get_audio(sdr_lib.CHAIN_WBFM)
        am = s.get_audio(sdr_lib.CHAIN_AM)
        cnt = s.counters()
    assert cnt["bytes_in"] == total and cnt["blocks_in"] == len(cuts)
    gold_spec, gold_frames = g.spectrum(iq)
    assert frames == gold_frames
    spec_check(spec, gold_spec)
    # audio produced so far = every output whose chunk is complete
    ga = g.wbfm(iq)
    n_fm = sdr_lib.wbfm_stream_audio_len(total)
    assert fm.size == n_fm and np.max(np.abs(fm - ga[:n_fm])) <= FM_AUDIO_ATOL
    gam = g.am(iq)
    n_am = (2 * (total // 2 // 200) + 2) // 3
    assert am.size == n_am and np.max(np.abs(am - gam[:n_am])) <= AM_AUDIO_ATOL


@pytest.mark.parametrize("chains", ["counter_only", "all_chains"])
@pytest.mark.parametrize("cuts_kind", ["baseline_blocks", "reference_512", "random"])
def test_streaming_counter_check(sdr_lib, g, cuts_kind, chains):
    """The firmware's own stream (test-mode counter) through process_samples: totals over the whole stream equal the
    golden check of the concatenated bytes however the stream is cut -- also when a break falls on a block
    boundary -- and the device stream buffer wraps several times on the way."""
    total = 262144 * 9 + 512 * 3
    u = (np.arange(total + 96) % 256).astype(np.uint8)
    u = np.delete(u, np.r_[70000:70032, 262144:262176, 262144 * 5 + 508:262144 * 5 + 540])[:total]  # three drops of 32 bytes
    u[262144 * 7 + 3] ^= 1                                                                         # one corrupted byte
    want = g.counter_check(u)
    assert want[0] == 5 and want[1] == 70000
    cuts = {"baseline_blocks": [262144] * 9 + [512 * 3], "reference_512": [512] * (total // 512),
            "random": random_cuts(total, 65536, 11)}[cuts_kind]
    mask = sdr_lib.CHAIN_COUNTER if chains == "counter_only" else (sdr_lib.CHAIN_COUNTER | sdr_lib.CHAIN_SPECTRUM |
                                                                   sdr_lib.CHAIN_WBFM | sdr_lib.CHAIN_AM)
    with sdr_lib.B200Sdr(chains=mask, slot_bytes=262144, ring_slots=4) as s:
        feed(s, u, cuts)
        assert s.get_counter_check() == want
        s.reset()                               # a new stream: totals and the carried byte start over
        feed(s, u[:4096], [4096])
        assert s.get_counter_check() == (0, None)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        with pytest.raises(sdr_lib.B200SdrError):
            s.get_counter_check()


def test_streaming_reset_starts_a_new_capture(sdr_lib, g):
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 91)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        s.process_samples(np.random.default_rng(1).integers(0, 256, 8192, dtype=np.uint8))
        s.reset()
        s.process_samples(iq)
        spec, frames = s.get_spectrum()
    assert frames == 255
    spec_check(spec, g.spectrum(iq)[0])


def test_streaming_ema_spectrum(sdr_lib, g):
    iq = g.synth(1, 262144 * 2, SYNTH_MULTITONE, 92)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM, avg_mode=AVG_EMA, ema_beta=0.1) as s:
        feed(s, iq, random_cuts(iq.size, 32768, 3))
        spec, _ = s.get_spectrum()
    spec_check_few_frames(spec, g.spectrum(iq, avg_mode=AVG_EMA, beta=0.1)[0])


# -------------------------------------------------------------------------------------- ingest
def test_ingest_bytes_match_reference_copy(sdr_lib, g, tmp_path):
    """Row a1-a7: the block that reaches the device is byte-identical to what the reference's own
    FSM + IRQ copy leaves in CommItf.buff for the same stream (oracle A), block by block."""
    ref = RefHost()
    data = g.synth(1, 512 * 6, SYNTH_COUNTER, 0)
    with sdr_lib.B200Sdr(slot_bytes=512, ring_slots=2, chains=sdr_lib.CHAIN_SPECTRUM) as s:
        for b in range(6):
            blk = data[512 * b:512 * (b + 1)]
            while s.process_samples(blk, allow_busy=True) == 1:
                s.sync()
            got, n = s.debug_last_block(512)
            assert n == 512 and np.array_equal(got, blk)
            if ref.available:
                dest, written = ref.read_packet(blk, 512)
                assert written == 512 and np.array_equal(dest[:512], got)
        assert s.counters() == {"bytes_in": 3072, "blocks_in": 6, "busy_returns": s.counters()["busy_returns"]}
    if ref.available:
        out, _ = ref.run_stream(data, 512, 7, str(tmp_path))
        assert np.array_equal(out, data)


def test_small_blocks_are_coalesced_into_ring_slots(sdr_lib, g):
    """The firmware hands over 512-byte URBs (usbh_rtlsdr.c:230).  By default they are appended to the
    open pinned slot and go to the device one full slot at a time; the getters submit what is
    pending, so nothing is ever missing from a result, and the last block is where it should be."""
    iq = g.synth(1, 262144 + 512 * 3, SYNTH_MULTITONE, 94)
    launches = {}
    for mode, submit in (("coalesced", 0), ("per_block", 4)):
        with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM, submit_bytes=submit) as s:
            before = s.kernel_launches()
            feed(s, iq, [512] * (iq.size // 512))
            got, n = s.debug_last_block(512)
            assert n == 512 and np.array_equal(got, iq[-512:])
            spec, frames = s.get_spectrum()
            launches[mode] = s.kernel_launches() - before
            assert s.counters()["blocks_in"] == iq.size // 512 and s.counters()["bytes_in"] == iq.size
        assert frames == (iq.size - 2048) // 1024 + 1
        spec_check(spec, g.spectrum(iq)[0])
    # per block: one k_spectrum launch per completed frame (every second 512-byte block); the streaming form folds the
    # finalize into the kernel's last CTA, so there is no second launch per slot
    assert launches["coalesced"] <= 4 and launches["per_block"] >= 250
    with pytest.raises(sdr_lib.B200SdrError) as ei:
        sdr_lib.B200Sdr(slot_bytes=4096, submit_bytes=8192)
    assert ei.value.status == sdr_lib.NOT_SUPPORTED


def test_audio_fifo_partial_pops_and_back_pressure(sdr_lib, g):
    """A small audio FIFO drained in odd-sized pieces: BUSY when a block's audio would not fit, nothing
    lost or reordered across partial pops and the FIFO's internal compaction."""
    total = 65536 * 24
    iq = g.synth(1, total, SYNTH_WBFM, 96)
    got_fm, got_am, pos, busy = [], [], 0, 0
    rng = np.random.default_rng(5)
    with sdr_lib.B200Sdr(slot_bytes=65536, ring_slots=3, audio_capacity=4096, submit_bytes=4,
                         chains=sdr_lib.CHAIN_WBFM | sdr_lib.CHAIN_AM) as s:
        while pos < total:
            rc = s.process_samples(iq[pos:pos + 65536], allow_busy=True)
            if rc == 1:
                busy += 1
                got_fm.append(s.get_audio(sdr_lib.CHAIN_WBFM, int(rng.integers(1, 1500))))
                got_am.append(s.get_audio(sdr_lib.CHAIN_AM, int(rng.integers(1, 300))))
                continue
            pos += 65536
            if rng.integers(0, 3) == 0:
                got_fm.append(s.get_audio(sdr_lib.CHAIN_WBFM, int(rng.integers(1, 700))))
        while True:
            a, b = s.get_audio(sdr_lib.CHAIN_WBFM, 333), s.get_audio(sdr_lib.CHAIN_AM, 77)
            got_fm.append(a)
            got_am.append(b)
            if a.size == 0 and b.size == 0:
                break
    fm, am = np.concatenate(got_fm), np.concatenate(got_am)
    assert busy > 0
    n_fm = sdr_lib.wbfm_stream_audio_len(total)
    n_am = (2 * (total // 2 // 200) + 2) // 3
    assert fm.size == n_fm and np.max(np.abs(fm - g.wbfm(iq)[:n_fm])) <= FM_AUDIO_ATOL
    assert am.size == n_am and np.max(np.abs(am - g.am(iq)[:n_am])) <= AM_AUDIO_ATOL


def test_ring_acquire_after_pending_blocks_keeps_stream_order(sdr_lib, g):
    iq = g.synth(1, 262144 + 4096, SYNTH_MULTITONE, 95)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        s.process_samples(iq[:4096])           # stays pending in the open slot
        slot = s.ring_acquire()                # must submit the pending bytes first
        slot[:] = iq[4096:]
        s.ring_commit(262144)
        spec, frames = s.get_spectrum()
    assert frames == (iq.size - 2048) // 1024 + 1
    spec_check(spec, g.spectrum(iq)[0])


def test_ring_acquire_commit_zero_copy(sdr_lib, g):
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 93)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        slot = s.ring_acquire()
        assert slot is not None and slot.size == 262144
        slot[:] = iq
        s.ring_commit(iq.size)
        spec, frames = s.get_spectrum()
    assert frames == 255
    spec_check(spec, g.spectrum(iq)[0])


def test_error_behaviour(sdr_lib):
    with sdr_lib.B200Sdr(slot_bytes=4096, ring_slots=2) as s:
        assert s.process_samples_raw(np.zeros(6, np.uint8)) == sdr_lib.NOT_SUPPORTED     # multiple-of-4 rule
        assert s.process_samples_raw(np.zeros(8192, np.uint8)) == sdr_lib.NOT_SUPPORTED  # larger than a slot
        assert s.process_samples_raw(np.zeros(0, np.uint8)) == sdr_lib.OK
        assert s.process_samples_raw(np.zeros(4096, np.uint8)) == sdr_lib.OK
    with pytest.raises(sdr_lib.B200SdrError) as ei:
        sdr_lib.B200Sdr(slot_bytes=510)
    assert ei.value.status == sdr_lib.NOT_SUPPORTED
    with pytest.raises(sdr_lib.B200SdrError):
        sdr_lib.B200Sdr(ring_slots=1)


def test_busy_when_ring_is_full(sdr_lib):
    """USBH_BUSY semantics: with every slot in flight the call returns BUSY and must be retried."""
    blk = np.zeros(1 << 22, np.uint8)
    with sdr_lib.B200Sdr(slot_bytes=1 << 22, ring_slots=2, chains=sdr_lib.CHAIN_SPECTRUM) as s:
        codes = [s.process_samples(blk, allow_busy=True) for _ in range(64)]
        s.sync()
        assert set(codes) <= {0, 1}
        assert s.counters()["blocks_in"] == codes.count(0)
        assert s.counters()["busy_returns"] == codes.count(1)


# ------------------------------------------------------------------- device generator + full size
@pytest.mark.parametrize("kind", [SYNTH_COUNTER, SYNTH_MULTITONE, SYNTH_WBFM, SYNTH_AM])
def test_device_generator_bit_identical(sdr, g, kind):
    n_cap, len_each = 3, 65536
    d = sdr.dev_alloc(n_cap * len_each)
    try:
        sdr.synth_fill_dev(d, n_cap, len_each, kind, first_capture=17)
        sdr.sync()
        got = sdr.to_host(d, n_cap * len_each)
    finally:
        sdr.dev_free(d)
    assert np.array_equal(got, g.synth(n_cap, len_each, kind, first_capture=17))


def test_full_size_capture_batch(sdr, g):
    """BASELINE configs[1..3] at full size: 10 s captures (48 MB each) generated on the device,
    run resident in HBM; captures picked from the batch are re-generated on the CPU and checked
    against the golden model; equal-seed captures give bitwise equal outputs."""
    n_cap, len_each = 6, 48_000_000
    d_iq = sdr.dev_alloc(n_cap * len_each)
    na, nam = 480000, 80000
    d_spec, d_fm, d_am = sdr.dev_alloc(4 * 1024 * n_cap), sdr.dev_alloc(4 * na * n_cap), sdr.dev_alloc(4 * nam * n_cap)
    try:
        # captures 0,1: multitone; 2,3: WBFM; 4,5: AM -- and capture 1/3/5 repeat seeds of 0/2/4
        for c, kind, seed in [(0, SYNTH_MULTITONE, 1000), (1, SYNTH_MULTITONE, 1000), (2, SYNTH_WBFM, 1001),
                              (3, SYNTH_WBFM, 1001), (4, SYNTH_AM, 1002), (5, SYNTH_AM, 1002)]:
            sdr.synth_fill_dev(d_iq + c * len_each, 1, len_each, kind, first_capture=seed)
        sdr.batch_spectrum_dev(d_iq, n_cap, len_each, d_spec)
        sdr.batch_wbfm_dev(d_iq, n_cap, len_each, d_fm)
        sdr.batch_am_dev(d_iq, n_cap, len_each, d_am)
        sdr.sync()
        spec = sdr.to_host(d_spec, 4 * 1024 * n_cap, np.float32).reshape(n_cap, 1024)
        fm = sdr.to_host(d_fm, 4 * na * n_cap, np.float32).reshape(n_cap, na)
        am = sdr.to_host(d_am, 4 * nam * n_cap, np.float32).reshape(n_cap, nam)
    finally:
        for p in (d_iq, d_spec, d_fm, d_am):
            sdr.dev_free(p)
    for a in (spec, fm, am):
        assert np.array_equal(a[0], a[1]) and np.array_equal(a[2], a[3]) and np.array_equal(a[4], a[5])
    gold, frames = g.spectrum(g.synth(1, len_each, SYNTH_MULTITONE, 1000))
    assert frames == 46874
    spec_check(spec[0], gold)
    assert np.max(np.abs(fm[2] - g.wbfm(g.synth(1, len_each, SYNTH_WBFM, 1001)))) <= FM_AUDIO_ATOL
    assert np.max(np.abs(am[4] - g.am(g.synth(1, len_each, SYNTH_AM, 1002)))) <= AM_AUDIO_ATOL


def test_kernel_launch_counter_moves(sdr, g):
    before = sdr.kernel_launches()
    sdr.spectrum(g.synth(1, 262144, SYNTH_MULTITONE, 0))
    assert sdr.kernel_launches() >= before + 1  # one 256 KiB block: k_spectrum with the finalize folded into its last CTA


# ------------------------------------------------------------------------------- more edge cases
def test_empty_and_misaligned_batches(sdr, sdr_lib):
    d = sdr.dev_alloc(4096)
    try:
        lib, ctx = sdr.lib, sdr.ctx
        assert lib.b200sdr_batch_spectrum_dev(ctx, d, 0, 2048, d) == sdr_lib.OK          # no captures
        assert lib.b200sdr_batch_wbfm_dev(ctx, d, 1, 0, d, None) == sdr_lib.OK           # empty capture
        assert lib.b200sdr_batch_am_dev(ctx, d, 0, 0, d) == sdr_lib.OK
        assert lib.b200sdr_batch_wbfm_dev(ctx, d, 1, 2044, d, None) == sdr_lib.NOT_SUPPORTED  # not 16-byte granular
        assert lib.b200sdr_batch_am_dev(ctx, d + 4, 1, 2048, d) == sdr_lib.NOT_SUPPORTED      # misaligned pointer
        assert lib.b200sdr_batch_spectrum_dev(ctx, d, 1, 2046, d) == sdr_lib.NOT_SUPPORTED    # multiple-of-4 rule
    finally:
        sdr.dev_free(d)


def test_many_small_captures(sdr, g):
    n_cap, len_each = 300, 4096
    iq = g.synth(n_cap, len_each, SYNTH_WBFM, 500)
    spec = sdr.spectrum(iq, n_cap)
    fm = sdr.wbfm(iq, n_cap)
    am = sdr.am(iq, n_cap)
    for c in (0, 1, 149, 299):
        blk = iq[c * len_each:(c + 1) * len_each]
        spec_check_few_frames(spec[c], g.spectrum(blk)[0])
        assert np.max(np.abs(fm[c] - g.wbfm(blk))) <= FM_AUDIO_ATOL
        assert np.max(np.abs(am[c] - g.am(blk))) <= AM_AUDIO_ATOL


@pytest.mark.parametrize("submit_bytes", [0, 4, 64], ids=["coalesced", "per_block", "every_64B"])
def test_streaming_tiny_blocks(sdr_lib, g, submit_bytes):
    """Blocks far smaller than a frame or a FIR chunk (down to the 4-byte minimum)."""
    total = 4 * 1500
    iq = g.synth(1, total, SYNTH_WBFM, 77)
    cuts = [4] * 300 + [8] * 100 + [12] * 50 + [4000 - 0]
    cuts = cuts[:-1] + [total - sum(cuts[:-1])]
    with sdr_lib.B200Sdr(slot_bytes=4096, ring_slots=3, submit_bytes=submit_bytes) as s:
        feed(s, iq, cuts)
        spec, frames = s.get_spectrum()
        fm = s.get_audio(sdr_lib.CHAIN_WBFM)
    gold, gframes = g.spectrum(iq)
    assert frames == gframes == 4
    spec_check_few_frames(spec, gold)
    n_fm = sdr_lib.wbfm_stream_audio_len(total)
    assert fm.size == n_fm and np.max(np.abs(fm - g.wbfm(iq)[:n_fm])) <= FM_AUDIO_ATOL


@pytest.mark.parametrize("submit_bytes", [0, 4], ids=["coalesced", "per_block"])
def test_streaming_wraps_the_device_stream_buffer_many_times(sdr_lib, g, submit_bytes):
    """The device side of the ring is one linear buffer of ring_slots x slot_bytes (+ slack); with two
    4 KiB slots a 400 KB stream wraps it about twenty times, each time carrying the unread tail of the
    slowest chain to the front.  Results must not notice."""
    total = 4 * 100_000
    iq = g.synth(1, total, SYNTH_WBFM, 78)
    with sdr_lib.B200Sdr(slot_bytes=4096, ring_slots=2, submit_bytes=submit_bytes) as s:
        fm, am = [], []
        pos = 0
        for n in random_cuts(total, 4096, 11):
            feed(s, iq[pos:pos + n], [n])
            pos += n
            if len(fm) * 7 % 5 == 0:
                fm.append(s.get_audio(sdr_lib.CHAIN_WBFM))
        fm.append(s.get_audio(sdr_lib.CHAIN_WBFM))
        am.append(s.get_audio(sdr_lib.CHAIN_AM))
        spec, frames = s.get_spectrum()
        got, n = s.debug_last_block(16)
        assert n > 0 and np.array_equal(got[:min(n, 16)], iq[total - n:total - n + min(n, 16)])
    fm, am = np.concatenate(fm), np.concatenate(am)
    gold, gframes = g.spectrum(iq)
    assert frames == gframes
    spec_check(spec, gold)
    n_fm = sdr_lib.wbfm_stream_audio_len(total)
    n_am = (2 * (total // 2 // 200) + 2) // 3
    assert fm.size == n_fm and np.max(np.abs(fm - g.wbfm(iq)[:n_fm])) <= FM_AUDIO_ATOL
    assert am.size == n_am and np.max(np.abs(am - g.am(iq)[:n_am])) <= AM_AUDIO_ATOL


def test_example_host_driver_runs(sdr_lib, g, tmp_path):
    """The C host driver (examples/host_driver.c): reference START/WAIT/COMPLETE cadence over the
    pinned ring; its audio files must equal the golden chains."""
    import subprocess
    exe = tmp_path / "host_driver"
    libdir = os.path.dirname(sdr_lib.LIB_PATH)
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "host_driver.c"),
                    "-L", libdir, "-lb200sdr", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    total = 2_400_000
    res = subprocess.run([str(exe), "--synthetic", "wbfm", str(total), "65024"], capture_output=True, text=True, cwd=tmp_path, check=True)
    assert f"{total} bytes in" in res.stdout and "frames averaged" in res.stdout
    iq = g.synth(1, total, SYNTH_WBFM, 0)
    fm = np.fromfile(tmp_path / "fm48k.f32", np.float32)
    am = np.fromfile(tmp_path / "am8k.f32", np.float32)
    n_fm = sdr_lib.wbfm_stream_audio_len(total)
    n_am = (2 * (total // 2 // 200) + 2) // 3
    assert fm.size == n_fm and np.max(np.abs(fm - g.wbfm(iq)[:n_fm])) <= FM_AUDIO_ATOL
    assert am.size == n_am and np.max(np.abs(am - g.am(iq)[:n_am])) <= AM_AUDIO_ATOL
    assert res.stdout.count("  bin ") == 5   # the five strongest spectrum bins are listed


def test_example_firmware_cadence_runs(sdr_lib, tmp_path):
    """examples/firmware_cadence.c: the firmware's one-buffer 512-byte superloop with process_samples at
    the COMPLETE point; every block must arrive (frame count) in both submit modes."""
    import subprocess
    exe = tmp_path / "firmware_cadence"
    libdir = os.path.dirname(sdr_lib.LIB_PATH)
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "firmware_cadence.c"),
                    "-L", libdir, "-lb200sdr", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    for total, submit in ((4_800_000, 0), (512_000, 4)):
        res = subprocess.run([str(exe), str(total), "512", str(submit), "262144", "1000"], capture_output=True, text=True, check=True)
        kv = dict(t.split("=") for t in res.stdout.split() if "=" in t)
        assert res.stdout.startswith("FW_CADENCE") and int(kv["blocks"]) == total // 512
        assert int(kv["frames"]) == (total - 2048) // 1024 + 1
        assert float(kv["realtime"]) > 1.0   # keeps up with a 2.4 MS/s dongle


def test_contexts_are_independent(sdr_lib, g):
    """Several dongles on one GPU = several contexts: blocks fed alternately, each context keeps its
    own carried state and audio FIFO (one USBH host handle per dongle in the reference)."""
    n_ctx, total, blk = 3, 65536 * 5 + 1000, 65536
    streams = [g.synth(1, total, kind, 200 + i) for i, kind in enumerate((SYNTH_WBFM, SYNTH_AM, SYNTH_MULTITONE))]
    ctxs = [sdr_lib.B200Sdr(slot_bytes=blk, ring_slots=3, submit_bytes=(0, 4, 8192)[i]) for i in range(n_ctx)]
    try:
        for pos in range(0, total, blk):
            for s, iq in zip(ctxs, streams):
                feed(s, iq[pos:pos + blk], [min(blk, total - pos)])
        for s, iq in zip(ctxs, streams):
            spec, frames = s.get_spectrum()
            gold, gframes = g.spectrum(iq)
            assert frames == gframes
            spec_check(spec, gold)
            fm, am = s.get_audio(sdr_lib.CHAIN_WBFM), s.get_audio(sdr_lib.CHAIN_AM)
            n_fm = sdr_lib.wbfm_stream_audio_len(total)
            n_am = (2 * (total // 2 // 200) + 2) // 3
            assert fm.size == n_fm and np.max(np.abs(fm - g.wbfm(iq)[:n_fm])) <= FM_AUDIO_ATOL
            assert am.size == n_am and np.max(np.abs(am - g.am(iq)[:n_am])) <= AM_AUDIO_ATOL
    finally:
        for s in ctxs:
            s.close()


# ------------------------------------------------------------- multi-GPU sharding
def _run_shards(pkg, devices, iq, n_cap, len_each, shards):
    """Every shard (lo, hi) of the batch on its own context (round-robin over `devices`); per-capture
    spectra / WBFM audio / AM audio gathered in capture order."""
    spec, fm, am = [], [], []
    ctxs = [pkg.B200Sdr(device=d) for d in devices]
    try:
        for i, (lo, hi) in enumerate(shards):
            if hi <= lo:
                continue
            s = ctxs[i % len(ctxs)]
            part = iq[lo * len_each:hi * len_each]
            spec.append(s.spectrum(part, hi - lo))
            fm.append(s.wbfm(part, hi - lo))
            am.append(s.am(part, hi - lo))
    finally:
        for s in ctxs:
            s.close()
    return np.concatenate(spec), np.concatenate(fm), np.concatenate(am)


@pytest.mark.parametrize("len_each", [262144, 16 * 190_000 + 16], ids=["block_256KiB", "3MB_capture"])
def test_sharded_batch_is_bitwise_the_single_context_batch(sdr_lib, g, len_each):
    """configs[4] shards 4096 captures over 1/2/4/8 GPUs.  A capture's results must not depend on which
    captures share its launch or on which GPU runs it: the same batch as ONE launch, as 2 / 4 / 8
    contiguous shards (stm32f7-rtlsdr_b200/sharding.py, the partition bench.py uses) and one capture at a time --
    on two devices when the box has them, else on two contexts of one -- gives bitwise equal spectra and audio.
    (The spectrum's summation tree follows from the capture length only, csrc/plan.h.)"""
    sharding = __import__("importlib").import_module("stm32f7-rtlsdr_b200.sharding")
    n_dev = 1
    try:
        import torch
        n_dev = max(1, torch.cuda.device_count())
    except Exception:
        pass
    n_cap = 16
    iq = np.concatenate([g.synth(1, len_each, (SYNTH_MULTITONE, SYNTH_WBFM, SYNTH_AM)[c % 3], 300 + c) for c in range(n_cap)])
    whole = _run_shards(sdr_lib, [0], iq, n_cap, len_each, [(0, n_cap)])
    gold, frames = g.spectrum(iq[:len_each])
    spec_check(whole[0][0], gold) if frames >= 200 else spec_check_few_frames(whole[0][0], gold)
    for world in (2, 4, 8, n_cap):
        devices = list(range(min(n_dev, 2))) if n_dev > 1 else [0, 0]
        got = _run_shards(sdr_lib, devices, iq, n_cap, len_each, sharding.all_shards(n_cap, world))
        for name, a, b in zip(("spectrum", "wbfm", "am"), whole, got):
            assert a.shape == b.shape and np.array_equal(a, b), f"{name}: {world} shards differ from the single batch"


def test_full_size_capture_alone_equals_in_batch(sdr, g):
    """Same at BASELINE's full size: a 10 s capture alone vs inside a batch of 5 (different neighbours)."""
    n_cap, len_each = 5, 48_000_000
    na = 480000
    d_iq = sdr.dev_alloc(n_cap * len_each)
    d_spec, d_fm = sdr.dev_alloc(4 * 1024 * (n_cap + 1)), sdr.dev_alloc(4 * na * (n_cap + 1))
    try:
        for c in range(n_cap):
            sdr.synth_fill_dev(d_iq + c * len_each, 1, len_each, SYNTH_WBFM if c % 2 else SYNTH_MULTITONE, first_capture=2000 + c)
        sdr.batch_spectrum_dev(d_iq, n_cap, len_each, d_spec)
        sdr.batch_wbfm_dev(d_iq, n_cap, len_each, d_fm)
        sdr.sync()
        spec = sdr.to_host(d_spec, 4 * 1024 * n_cap, np.float32).reshape(n_cap, 1024)
        fm = sdr.to_host(d_fm, 4 * na * n_cap, np.float32).reshape(n_cap, na)
        for c in (0, 3, 4):
            sdr.batch_spectrum_dev(d_iq + c * len_each, 1, len_each, d_spec + 4 * 1024 * n_cap)
            sdr.batch_wbfm_dev(d_iq + c * len_each, 1, len_each, d_fm + 4 * na * n_cap)
            sdr.sync()
            assert np.array_equal(sdr.to_host(d_spec + 4 * 1024 * n_cap, 4 * 1024, np.float32), spec[c])
            assert np.array_equal(sdr.to_host(d_fm + 4 * na * n_cap, 4 * na, np.float32), fm[c])
    finally:
        for p in (d_iq, d_spec, d_fm):
            sdr.dev_free(p)


def test_one_big_slot_fits_a_small_audio_fifo(sdr_lib, g):
    """slot_bytes = 1 MiB with audio_capacity = 4096 used to answer BUSY for ever (one slot makes
    10 486 WBFM samples); b200sdr_create now raises the FIFO to what one full slot produces."""
    blk = 1 << 20
    iq = g.synth(1, blk, SYNTH_WBFM, 5)
    with sdr_lib.B200Sdr(slot_bytes=blk, ring_slots=2, audio_capacity=4096) as s:
        assert s.process_samples(iq, allow_busy=True) == sdr_lib.OK
        fm = s.get_audio(sdr_lib.CHAIN_WBFM)
        n_fm = sdr_lib.wbfm_stream_audio_len(blk)
        assert fm.size == n_fm and np.max(np.abs(fm - g.wbfm(iq)[:n_fm])) <= FM_AUDIO_ATOL
        assert s.process_samples(iq, allow_busy=True) == sdr_lib.OK     # and again after the pop
