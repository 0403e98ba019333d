"""The product's kernel SOURCE (csrc/*.cuh) run on the host under tests/emu/cuda_emu.h and checked
against oracle B: validates the four-step FFT indexing, the scatter-FIR bookkeeping, the carried
state, the scans and the launch planning without a GPU.  The fp32x2 PTX, TMA and mbarrier parts
have host stand-ins, so this is a logic check -- the parity tests proper are the -m gpu tests."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_api import (AVG_EMA, SYNTH_AM, SYNTH_MULTITONE, SYNTH_WBFM, WIN_BLACKMAN, WIN_HANN, Golden, extreme_patterns,
                        wrap_phase)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    e = C.CDLL(os.path.join(ROOT, "tests", "emu", "libemu_kernels.so"))
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    e.emu_spectrum.argtypes = [vp, u32, u64, vp, u32, C.c_int, C.c_float, vp]
    e.emu_wbfm_batch.argtypes = [vp, u32, u64, u32, vp, vp]
    e.emu_wbfm_stream.argtypes = [vp, u32, u64, vp, vp, vp, vp, u32]
    e.emu_wbfm_tc_batch.argtypes = [vp, u32, u64, u32, vp, vp, vp, vp]
    e.emu_am_batch.argtypes = [vp, u32, u64, u32, vp, vp]
    e.emu_am_stream.argtypes = [vp, u32, u64, vp, vp, vp, vp, u32]
    return e


@pytest.fixture(scope="module")
def g():
    return Golden()


def padded(a):
    b = np.zeros(a.size + 64, np.uint8)
    b[: a.size] = a
    return b


def test_spectrum_launch_plan(emu):
    """csrc/plan.h plan_spectrum: every frame is covered exactly once; frames per warp (hence the summation
    tree, hence the bits of a capture's spectrum) depend on the capture LENGTH only -- not on the batch size,
    not on the device; the persistent grid never exceeds the resident CTA slots."""
    emu.emu_plan_spectrum.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]

    def plan(nbytes, ncap, sms=148):
        out = (C.c_uint32 * 4)()
        emu.emu_plan_spectrum(nbytes, ncap, sms, out)
        return tuple(out)

    for nbytes, ncap in ((48_000_000, 512), (48_000_000, 96), (48_000_000, 4), (48_000_000, 1), (262144, 1),
                         (2048, 1), (2048 + 1024 * 5, 3), (4 * 262144, 7), (48_000_000, 65535), (480_000_000, 2)):
        frames, fpw, units, grid = plan(nbytes, ncap)
        assert frames == (nbytes // 2 - 1024) // 512 + 1
        assert 1 <= fpw <= 64 and (units - 1) * fpw * 4 < frames <= units * fpw * 4   # no empty unit, all frames covered
        assert 1 <= grid <= min(units * ncap, 2 * 148)
        for other_ncap, sms in ((1, 148), (8, 148), (512, 148), (4096, 74), (3, 132)):
            assert plan(nbytes, other_ncap, sms)[:3] == (frames, fpw, units)           # batch-shape / device independent
    assert plan(2046, 1) == (0, 0, 0, 0)                   # shorter than one frame: nothing to launch
    assert plan(48_000_000, 1)[1:] == (40, 293, 293)       # one 10 s capture: one wave of 296 CTA slots
    assert plan(48_000_000, 512)[3] == 296                 # a batch: persistent grid, 293 x 512 units
    assert plan(262144, 1)[1:] == (1, 64, 64)              # one 256 KiB block: 64 CTAs of 4 single-frame warps
    assert plan(480_000_000, 1)[1] == 64                   # long captures: capped (partials < 2 % of the input)


def test_convert_and_generator_kernels(emu, g):
    """k_convert_cf32 (bit-exact, with and without a window) and k_synth (bytes identical to the host generator that
    the oracle shares through include/b200sdr_synth.h) under the host emulation."""
    emu.emu_convert_cf32.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    emu.emu_synth.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p]
    iq = np.random.default_rng(3).integers(0, 256, 16 * 1500 + 16 * 3, dtype=np.uint8)
    iq[:256] = np.arange(256)                       # every byte value
    out = np.full(iq.size, np.nan, np.float32)
    emu.emu_convert_cf32(iq.ctypes.data, iq.size, None, out.ctypes.data)
    assert np.array_equal(out.astype(np.float64), g.convert(iq))
    w = g.window(WIN_HANN).astype(np.float32)
    emu.emu_convert_cf32(iq.ctypes.data, iq.size, w.ctypes.data, out.ctypes.data)
    want = (g.convert(iq).astype(np.float32) * np.repeat(np.resize(w, iq.size // 2), 2)).astype(np.float32)
    assert np.array_equal(out, want)                # one rounded fp32 multiply by the fp32 window value
    lut = np.sin(2 * np.pi * (np.arange(4097) % 4096) / 4096).astype(np.float32)
    for kind in (SYNTH_MULTITONE, SYNTH_WBFM, SYNTH_AM, 0):
        n_cap, len_each = 2, 16 * 300
        buf = np.zeros(n_cap * len_each, np.uint8)
        emu.emu_synth(buf.ctypes.data, n_cap, len_each, kind, 5, lut.ctypes.data)
        assert np.array_equal(buf, g.synth(n_cap, len_each, kind, 5))


def aligned_bytes(n, offset):
    """n bytes starting `offset` bytes after a 64-byte boundary"""
    raw = np.zeros(n + 128, np.uint8)
    start = (-raw.ctypes.data) % 64 + offset
    return raw[start:start + n]


@pytest.mark.parametrize("nbytes,offset", [(4, 0), (12, 4), (16, 0), (20, 12), (2048 + 16, 0), (2048 + 32, 0), (2048 + 36, 8),
                                           (16384, 0), (16384 + 2048 + 20, 4), (3 * 16384 + 2064, 0)])
def test_counter_kernel_logic(emu, g, nbytes, offset):
    """k_counter_check under the host emulation: warp-contiguous fast path, the bounds-checked last warps, head words
    before the first 16-byte boundary, tail words, breaks on every kind of boundary, the streaming state."""
    emu.emu_counter_check.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.c_void_p, C.c_uint64]

    def run(u, expect=-1, state=None, pos_base=0):
        n, f = C.c_uint64(0), C.c_uint64(0)
        emu.emu_counter_check(u.ctypes.data, u.size, expect, C.byref(n), C.byref(f), state, pos_base)
        return n.value, (None if f.value == 2**64 - 1 else f.value)

    rng = np.random.default_rng(nbytes + offset)
    u = aligned_bytes(nbytes, offset)
    u[:] = (np.arange(nbytes) + 250) % 256
    assert run(u) == (0, None) and run(u, 250) == (0, None) and run(u, 3) == (1, 0)
    for _ in range(3):
        v = aligned_bytes(nbytes, offset)
        v[:] = u
        where = [w for w in list(rng.integers(0, nbytes, size=4)) + [nbytes - 1, 3, 4, 15, 16, 511, 512, 2047, 2048, 2063, 16383, 16384]
                 if w < nbytes]
        v[where] = rng.integers(0, 256, size=len(where), dtype=np.uint8)
        for expect in (-1, 250):
            assert run(v, expect) == g.counter_check(v, expect)
    # streaming: two consecutive blocks with a carried byte; totals carry absolute positions
    if nbytes >= 16:
        assert emu.emu_sizeof_counter_state() == 24
        st = np.zeros(3, np.uint64)
        st[1], st[2] = 2**64 - 1, 0xFFFFFFFF   # n_breaks = 0, first_break = none, expect = unknown
        cut = (nbytes // 2) & ~3
        a, b = aligned_bytes(cut, offset), aligned_bytes(nbytes - cut, (offset + cut) % 16)
        whole = u.copy()
        whole[cut] ^= 0x10                      # a break exactly on the block boundary (and the byte after it)
        a[:], b[:] = whole[:cut], whole[cut:]
        run(a, state=st.ctypes.data, pos_base=0)
        run(b, state=st.ctypes.data, pos_base=cut)
        want = g.counter_check(whole)
        assert (int(st[0]), int(st[1])) == (want[0], want[1]) and int(st[2]) & 0xFFFFFFFF == (int(whole[-1]) + 1) % 256


def test_counter_kernel_random_streams(emu, g):
    """property check (hypothesis): any length, any 4-byte start alignment, any set of corrupted bytes, any first-byte
    expectation -- the emulated kernel and the golden definition agree."""
    from hypothesis import given, settings, strategies as st

    emu.emu_counter_check.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.c_void_p, C.c_uint64]

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 1400), st.integers(0, 3), st.integers(0, 255), st.lists(st.integers(0, 5599), max_size=6),
           st.integers(-1, 255), st.integers(0, 2**32))
    def check(words, off_words, start, hits, expect, seed):
        n = 4 * words
        u = aligned_bytes(n, 4 * off_words)
        u[:] = (np.arange(n) + start) % 256
        rng = np.random.default_rng(seed)
        for h in hits:
            u[h % n] = rng.integers(0, 256)
        cnt, first = C.c_uint64(0), C.c_uint64(0)
        emu.emu_counter_check(u.ctypes.data, n, expect, C.byref(cnt), C.byref(first), None, 0)
        assert (cnt.value, None if first.value == 2**64 - 1 else first.value) == g.counter_check(u, expect)

    check()


@pytest.mark.parametrize("frames_per_warp,window", [(1, WIN_HANN), (3, WIN_HANN), (8, WIN_BLACKMAN)])
def test_spectrum_kernel_logic(emu, g, frames_per_warp, window):
    n = 512 * 21 + 1024 + 100
    iq = padded(g.synth(2, 2 * n, SYNTH_MULTITONE, 5))
    w = g.window(window).astype(np.float32)
    out = np.zeros(2048, np.float32)
    frames = emu.emu_spectrum(iq.ctypes.data, 2, 2 * n, w.ctypes.data, frames_per_warp, 0, 0.0, out.ctypes.data)
    assert frames == 22
    for c in range(2):
        gold, _ = g.spectrum(iq[c * 2 * n:(c + 1) * 2 * n], window=window)
        rel = np.abs(out[c * 1024:(c + 1) * 1024] - gold) / gold
        assert rel.max() < 1e-4 and rel.mean() < 5e-6  # host emulation has no FMA; the GPU bound is 1e-5


def test_spectrum_kernel_ema(emu, g):
    n = 512 * 12 + 1024
    iq = padded(g.synth(1, 2 * n, SYNTH_MULTITONE, 2))
    w = g.window(WIN_HANN).astype(np.float32)
    out = np.zeros(1024, np.float32)
    emu.emu_spectrum(iq.ctypes.data, 1, 2 * n, w.ctypes.data, 2, 1, 0.1, out.ctypes.data)
    gold, _ = g.spectrum(iq[: 2 * n], avg_mode=AVG_EMA, beta=0.1)
    assert np.max(np.abs(out - gold) / gold) < 1e-4


@pytest.mark.parametrize("n,ncap,tps", [(50000, 2, 0), (50000, 1, 1), (30720, 1, 2), (1208, 1, 0)])
def test_wbfm_kernel_logic(emu, g, n, ncap, tps):
    nb = 2 * n
    iq = padded(g.synth(ncap, nb, SYNTH_WBFM, 7))
    m1 = -(-n // 10)
    m2 = -(-m1 // 5)
    audio = np.full(m2 * ncap, np.nan, np.float32)
    disc = np.full(m1 * ncap, np.nan, np.float32)
    emu.emu_wbfm_batch(iq.ctypes.data, ncap, nb, tps, audio.ctypes.data, disc.ctypes.data)
    for c in range(ncap):
        ga, gd = g.wbfm(iq[c * nb:(c + 1) * nb], want_disc=True)
        assert np.max(np.abs(wrap_phase(disc[c * m1:(c + 1) * m1] - gd))) < 1e-5
        assert np.max(np.abs(audio[c * m2:(c + 1) * m2] - ga)) < 1e-5


@pytest.mark.parametrize("n,ncap,tps", [(160 * 125 * 3, 1, 0), (160 * 125 * 2 + 1232, 2, 1), (20000 + 160 * 7 + 8, 1, 2),
                                        (8, 1, 0), (160 * 125, 3, 0), (50008, 1, 1)])
def test_wbfm_tensor_engine_logic(emu, g, n, ncap, tps):
    """csrc/wbfm_tc.cuh under the host emulation: the banded-Toeplitz product (three signed 8-bit tap slices on the raw
    bytes, exact integers -- the part the tensor cores do on the device), the 128B-swizzled B image, the slice combine
    with the offset removed exactly, the first row of a capture (zero-filled history), ragged ends, segments with
    pre-roll, several work items per CTA; against oracle B."""
    nb = 2 * n
    iq = padded(g.synth(ncap, nb, SYNTH_WBFM, 11))
    m1 = -(-n // 10)
    m2 = -(-m1 // 5)
    audio = np.full(m2 * ncap, np.nan, np.float32)
    disc = np.full(m1 * ncap, np.nan, np.float32)
    acc = np.zeros((128, 112), np.int32)
    q = np.zeros((3, 80), np.int8)
    e = emu.emu_wbfm_tc_batch(iq.ctypes.data, ncap, nb, tps, audio.ctypes.data, disc.ctypes.data, acc.ctypes.data, q.ctypes.data)
    # the slices reproduce the float64 taps to 2^-22 / 2^e, i.e. 3.6e-7 of the largest tap
    h = g.taps(0)
    hq = (q[0] * 2.0**-7 + q[1] * 2.0**-14 + q[2] * 2.0**-21) / 2.0**e
    assert np.max(np.abs(hq - h)) <= 2.0**-22 / 2.0**e * 1.001 and np.max(np.abs(hq - h)) < 4e-7 * np.max(np.abs(h))
    # raw accumulators of tile 0 = the integer FIR of the raw bytes, slice by slice (row 0: zero-filled history)
    u = iq[:nb].astype(np.int64)
    # (column 34 s + 2 j + c holds output o = j - 1 of the row)
    for r in (0, 1, 7, 124):
        for j in (0, 1, 8, 9, 16):
            m = 16 * r - 1 + j
            if 10 * m >= n or m < 0:
                continue
            for c in (0, 1):
                for s in range(3):
                    want = sum(int(q[s][t]) * int(u[2 * (10 * m - t) + c]) for t in range(80) if 0 <= 10 * m - t < n)
                    assert acc[r, 34 * s + 2 * j + c] == want, (r, j, c, s)
    for c in range(ncap):
        ga, gd = g.wbfm(iq[c * nb:(c + 1) * nb], want_disc=True)
        derr = np.abs(wrap_phase(disc[c * m1:(c + 1) * m1] - gd))
        # outputs 1..7 are the filter's rise from nothing through its outermost taps (1e-3 of the largest and smaller),
        # where the 2^-22 tap quantisation is 1e-4 of the tap: phase within 1e-4 rad there, 1e-5 everywhere else
        assert np.max(derr[:8]) < 1e-4 and (derr.size <= 8 or np.max(derr[8:]) < 1e-5)
        assert np.max(np.abs(audio[c * m2:(c + 1) * m2] - ga)) < 1e-5


@pytest.mark.parametrize("name", ["fullscale_tone", "half_lsb_rotation", "tone_on_dc"])
def test_wbfm_tensor_engine_extreme_bytes(emu, g, name):
    """the smallest and the largest phasors a u8 stream can carry: the offset cancels in exact integer / half-integer
    arithmetic here, so the bound is the north_star's 1e-4 rad with room to spare"""
    n = 30720 + 1208
    nb = 2 * n
    const, rot = extreme_patterns(nb)
    iq = padded(rot[name])
    m1 = -(-n // 10)
    m2 = -(-m1 // 5)
    audio = np.full(m2, np.nan, np.float32)
    disc = np.full(m1, np.nan, np.float32)
    emu.emu_wbfm_tc_batch(iq.ctypes.data, 1, nb, 1, audio.ctypes.data, disc.ctypes.data, None, None)
    ga, gd = g.wbfm(iq[:nb], want_disc=True)
    # the first outputs of a capture are the filter's rise from nothing (|y| ~ 1e-6): skip what the fp32 engine's test skips
    derr, aerr = np.max(np.abs(wrap_phase(disc - gd))[16:]), np.max(np.abs(audio - ga)[8:])
    print(name, "disc err", derr, "audio err", aerr)
    assert derr < 2e-5 and aerr < 2e-5


@pytest.mark.parametrize("tps", [0, 1, 2])
def test_wbfm_streaming_state_carry(emu, g, tps):
    """tps = tiles per segment of every streaming launch (0: one CTA walks the whole block; otherwise
    the block is split over CTAs: segment 0 continues the carried state, the others pre-roll a tile)"""
    nch = 700
    ch = emu.emu_fm_chunk()                     # input samples per thread-chunk (csrc/wbfm.cuh B200_FM_CHUNK)
    opt = ch // 10
    n = ch * nch
    iq = padded(g.synth(1, 2 * n, SYNTH_WBFM, 3))
    ga, gd = g.wbfm(iq[: 2 * n], want_disc=True)
    state = np.zeros(emu.emu_sizeof_fm_state(), np.uint8)
    rng = np.random.default_rng(1)
    pos, outa, outd = 0, [], []
    small = [1, 1, 2, 1, 3, 4, 5, 1]  # launches shorter than the 49-sample history of stage 2
    while pos < nch:
        k = small.pop(0) if small else int(min(nch - pos, rng.integers(1, 300)))
        k = int(min(k, nch - pos))
        a = np.zeros(k * opt // 5 + 2, np.float32)
        d = np.zeros(k * opt, np.float32)
        na = C.c_uint32(0)
        emu.emu_wbfm_stream(iq[pos * 2 * ch:].ctypes.data, k, pos, state.ctypes.data, a.ctypes.data, C.byref(na), d.ctypes.data, tps)
        outa.append(a[: na.value])
        outd.append(d)
        pos += k
    outa, outd = np.concatenate(outa), np.concatenate(outd)
    assert outa.size == ga.size
    assert np.max(np.abs(outa - ga)) < 1e-5 and np.max(np.abs(wrap_phase(outd - gd))) < 1e-5


@pytest.mark.parametrize("name", ["all0", "all255", "half_lsb", "split", "square", "fullscale_tone", "half_lsb_rotation", "tone_on_dc"])
def test_fir_kernels_extreme_bytes(emu, g, name):
    n = 30720 + 1208
    nb = 2 * n
    const, rot = extreme_patterns(nb)
    iq = padded(rot[name] if name in rot else const[name])
    if name in rot:
        m1 = -(-n // 10)
        m2 = -(-m1 // 5)
        audio = np.full(m2, np.nan, np.float32)
        disc = np.full(m1, np.nan, np.float32)
        emu.emu_wbfm_batch(iq.ctypes.data, 1, nb, 1, audio.ctypes.data, disc.ctypes.data)
        ga, gd = g.wbfm(iq[:nb], want_disc=True)
        # the first outputs see the filter's rise from x[n < 0] = 0, where |y| passes through ~0.
        # The offset cancels in the accumulators (partial sums within +-0.5 of full scale), so the error is ~2e-7
        # of FULL SCALE whatever the signal level: < 1e-6 rad for ordinary signals, 6.5e-5 rad at worst for the
        # half-LSB phasor (|y| = 0.004), the smallest rotating signal a u8 stream can carry.
        tol = 1e-4 if name == "half_lsb_rotation" else 2e-6
        assert np.max(np.abs(wrap_phase(disc - gd))[16:]) < tol and np.max(np.abs(audio - ga)[8:]) < tol
    m3 = g.lib.gold_am_audio_len(n)
    a = np.full(m3, np.nan, np.float32)
    emu.emu_am_batch(iq.ctypes.data, 1, nb, 1, a.ctypes.data, None)
    assert np.max(np.abs(a - g.am(iq[:nb]))) < 2e-6


@pytest.mark.parametrize("n,ncap", [(200 * 128 * 3 + 40, 2), (1208, 1)])
def test_am_kernel_logic_and_segment_invariance(emu, g, n, ncap):
    nb = 2 * n
    iq = padded(g.synth(ncap, nb, SYNTH_AM, 7))
    m3 = g.lib.gold_am_audio_len(n)
    ref = None
    for tps in (0, 1, 2):
        audio = np.full(m3 * ncap, np.nan, np.float32)
        emu.emu_am_batch(iq.ctypes.data, ncap, nb, tps, audio.ctypes.data, None)
        if ref is None:
            ref = audio.copy()
        assert np.array_equal(audio, ref)  # FIR-only front end: bitwise independent of the segment split
        for c in range(ncap):
            assert np.max(np.abs(audio[c * m3:(c + 1) * m3] - g.am(iq[c * nb:(c + 1) * nb]))) < 1e-6


@pytest.mark.parametrize("tps", [0, 1, 2])
def test_am_streaming_state_carry(emu, g, tps):
    nch = 500
    n = 200 * nch
    iq = padded(g.synth(1, 2 * n, SYNTH_AM, 3))
    ga = g.am(iq[: 2 * n])
    fs = np.zeros(emu.emu_sizeof_am_front_state(), np.uint8)
    bs = np.zeros(emu.emu_sizeof_am_back_state(), np.uint8)
    rng = np.random.default_rng(1)
    pos, outa = 0, []
    small = [1, 1, 2, 1, 3, 11, 12, 13, 1, 5]  # launches shorter than the 12-output reach of stage 2
    while pos < nch:
        k = small.pop(0) if small else int(min(nch - pos, rng.integers(1, 300)))
        k = int(min(k, nch - pos))
        a = np.zeros(k + 2, np.float32)
        na = C.c_uint32(0)
        emu.emu_am_stream(iq[pos * 400:].ctypes.data, k, pos, fs.ctypes.data, bs.ctypes.data, a.ctypes.data, C.byref(na), tps)
        outa.append(a[: na.value])
        pos += k
    outa = np.concatenate(outa)
    assert outa.size == ga.size and np.max(np.abs(outa - ga)) < 1e-6
