"""ctypes view of the oracles for the tests (checker only -- never imported by the product)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")

WIN_RECT, WIN_HANN, WIN_BLACKMAN = 0, 1, 2
AVG_MEAN, AVG_EMA = 0, 1
SYNTH_COUNTER, SYNTH_MULTITONE, SYNTH_WBFM, SYNTH_AM = 0, 1, 2, 3


def _load(name):
    lib = C.CDLL(os.path.join(ORACLE, name))
    vp, sz, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64
    lib.gold_sizeof_real.restype = C.c_int
    lib.gold_ingest_copy.restype = sz
    lib.gold_ingest_copy.argtypes = [vp, vp, C.c_uint16]
    lib.gold_convert.argtypes = [vp, sz, vp]
    lib.gold_counter_check.restype = u64
    lib.gold_counter_check.argtypes = [vp, sz, C.c_int, C.POINTER(u64)]
    lib.gold_convert_window.argtypes = [vp, sz, C.c_int, vp]
    lib.gold_window.argtypes = [C.c_int, C.c_int, vp]
    lib.gold_fft.argtypes = [vp, vp, C.c_int]
    lib.gold_dft_naive.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.gold_spectrum.restype = u64
    lib.gold_spectrum.argtypes = [vp, sz, C.c_int, C.c_int, C.c_double, vp]
    lib.gold_kaiser_lowpass.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, vp]
    lib.gold_taps.restype = C.c_int
    lib.gold_taps.argtypes = [C.c_int, vp]
    lib.gold_deemph_alpha.restype = C.c_double
    lib.gold_dcblock_rho.restype = C.c_double
    for f in ("gold_wbfm_disc_len", "gold_wbfm_audio_len", "gold_am_audio_len"):
        getattr(lib, f).restype = sz
        getattr(lib, f).argtypes = [sz]
    lib.gold_wbfm.argtypes = [vp, sz, vp, vp]
    lib.gold_wbfm_stage1.argtypes = [vp, sz, vp]
    lib.gold_am.argtypes = [vp, sz, vp]
    lib.gold_synth_fill.argtypes = [vp, u32, u64, u32, u64]
    for f in ("gold_time_spectrum", "gold_time_wbfm", "gold_time_am"):
        getattr(lib, f).restype = C.c_double
        getattr(lib, f).argtypes = [vp, sz, u32, u32, C.c_int]
    return lib


class Golden:
    def __init__(self, f32=False):
        self.lib = _load("libgolden_f32.so" if f32 else "libgolden.so")
        self.real = np.float32 if f32 else np.float64
        assert self.lib.gold_sizeof_real() == np.dtype(self.real).itemsize

    def synth(self, n_captures, len_each, kind, first_capture=0):
        buf = np.zeros(n_captures * len_each + 64, np.uint8)  # padded: kernels may read to a 16-byte boundary
        self.lib.gold_synth_fill(buf.ctypes.data, n_captures, len_each, kind, first_capture)
        return buf[: n_captures * len_each]

    def counter_check(self, iq, expect_first=-1):
        iq = np.ascontiguousarray(iq, np.uint8)
        first = C.c_uint64(0)
        n = self.lib.gold_counter_check(iq.ctypes.data, iq.size, expect_first, C.byref(first))
        return int(n), (None if first.value == 2**64 - 1 else int(first.value))

    def convert(self, iq, window=None):
        iq = np.ascontiguousarray(iq, np.uint8)
        out = np.zeros(iq.size, self.real)
        if window is None:
            self.lib.gold_convert(iq.ctypes.data, iq.size // 2, out.ctypes.data)
        else:
            self.lib.gold_convert_window(iq.ctypes.data, iq.size // 2, window, out.ctypes.data)
        return out

    def window(self, kind, n=1024):
        w = np.zeros(n, self.real)
        self.lib.gold_window(kind, n, w.ctypes.data)
        return w

    def fft(self, x):
        re = np.ascontiguousarray(x.real, self.real).copy()
        im = np.ascontiguousarray(x.imag, self.real).copy()
        self.lib.gold_fft(re.ctypes.data, im.ctypes.data, re.size)
        return re + 1j * im

    def dft_naive(self, x):
        re = np.ascontiguousarray(x.real, np.float64)
        im = np.ascontiguousarray(x.imag, np.float64)
        ore, oim = np.zeros_like(re), np.zeros_like(im)
        self.lib.gold_dft_naive(re.ctypes.data, im.ctypes.data, re.size, ore.ctypes.data, oim.ctypes.data)
        return ore + 1j * oim

    def spectrum(self, iq, window=WIN_HANN, avg_mode=AVG_MEAN, beta=0.0):
        iq = np.ascontiguousarray(iq, np.uint8)
        out = np.zeros(1024, self.real)
        frames = self.lib.gold_spectrum(iq.ctypes.data, iq.size // 2, window, avg_mode, beta, out.ctypes.data)
        return out, int(frames)

    def taps(self, which):
        h = np.zeros(256, np.float64)
        n = self.lib.gold_taps(which, h.ctypes.data)
        return h[:n].copy()

    def wbfm(self, iq, want_disc=False):
        iq = np.ascontiguousarray(iq, np.uint8)
        n = iq.size // 2
        audio = np.zeros(self.lib.gold_wbfm_audio_len(n), self.real)
        disc = np.zeros(self.lib.gold_wbfm_disc_len(n), self.real)
        self.lib.gold_wbfm(iq.ctypes.data, n, audio.ctypes.data, disc.ctypes.data)
        return (audio, disc) if want_disc else audio

    def am(self, iq):
        iq = np.ascontiguousarray(iq, np.uint8)
        n = iq.size // 2
        audio = np.zeros(self.lib.gold_am_audio_len(n), self.real)
        self.lib.gold_am(iq.ctypes.data, n, audio.ctypes.data)
        return audio

    def ingest_copy(self, src, length):
        src = np.ascontiguousarray(src, np.uint8)
        padded = np.zeros(length + 8, np.uint8)
        padded[: min(src.size, length + 4)] = src[: length + 4]
        dest = np.full(length + 8, 0xEE, np.uint8)
        written = self.lib.gold_ingest_copy(dest.ctypes.data, padded.ctypes.data, length)
        return dest, int(written)


class RefHost:
    """Oracle A: the reference's own unmodified C (oracle/_ref, built from /root/reference)."""

    def __init__(self):
        self.so = os.path.join(ORACLE, "_ref", "libref_ingest.so")
        self.cli = os.path.join(ORACLE, "_ref", "ref_ingest_cli")
        self.available = os.path.exists(self.so) and os.path.exists(self.cli)
        if self.available:
            self.lib = C.CDLL(self.so)
            self.lib.ref_read_packet.restype = C.c_uint32
            self.lib.ref_read_packet.argtypes = [C.c_void_p, C.c_void_p, C.c_uint16]

    def read_packet(self, src, length):
        src = np.ascontiguousarray(src, np.uint8)
        padded = np.zeros(length + 8, np.uint8)
        padded[: min(src.size, length + 4)] = src[: length + 4]
        dest = np.full(length + 8, 0xEE, np.uint8)
        written = self.lib.ref_read_packet(dest.ctypes.data, padded.ctypes.data, length)
        return dest, int(written)

    def run_stream(self, data, buff_size, tim_cnt, tmpdir):
        """Feed `data` through the reference class FSM + HCD IRQ path; returns (bytes, log text)."""
        inp, outp = os.path.join(tmpdir, "in.bin"), os.path.join(tmpdir, "out.bin")
        np.ascontiguousarray(data, np.uint8).tofile(inp)
        res = subprocess.run([self.cli, inp, str(buff_size), str(tim_cnt), outp], capture_output=True, text=True, check=True)
        return np.fromfile(outp, np.uint8), res.stdout


def wrap_phase(a):
    return (np.asarray(a) + np.pi) % (2 * np.pi) - np.pi


def extreme_patterns(nb):
    """byte patterns that stress the raw-byte FIR form (cplx2.cuh form C).  Constant ones (AM only: a
    constant input puts the FM discriminator on its branch cut): the largest offsets the accumulators ever
    cancel (all 0 / all 255), the smallest signal a u8 stream can carry (127 / 128: u - 127.5 = -+0.5), both
    halves of the subnormal / normal split side by side, a full-scale square wave.  Rotating ones (FM and
    AM): a full-scale tone (bytes 0 and 255 occur), the smallest rotating phasor a u8 stream can carry
    (I, Q in {127, 128}), and a small tone on top of a large DC offset."""
    n = nb // 2
    const, rot = {}, {}
    for name, (a, b) in {"all0": (0, 0), "all255": (255, 255), "half_lsb": (127, 128), "split": (127, 255)}.items():
        v = np.empty(nb, np.uint8)
        v[0::2], v[1::2] = a, b
        const[name] = v
    sq = np.zeros(nb, np.uint8)
    sq[(np.arange(nb) // 14) % 2 == 0] = 255
    const["square"] = sq
    ph = 2 * np.pi * 50e3 / 2.4e6 * np.arange(n)
    def pack(i, q):
        v = np.empty(nb, np.uint8)
        v[0::2], v[1::2] = np.clip(np.rint(i), 0, 255), np.clip(np.rint(q), 0, 255)
        return v
    rot["fullscale_tone"] = pack(127.5 + 127.5 * np.cos(ph), 127.5 + 127.5 * np.sin(ph))
    rot["half_lsb_rotation"] = pack(127.5 + 0.5 * np.sign(np.cos(ph + 0.3)), 127.5 + 0.5 * np.sign(np.sin(ph + 0.3)))
    rot["tone_on_dc"] = pack(227.5 + 20 * np.cos(ph), 27.5 + 20 * np.sin(ph))
    return const, rot


