"""ctypes binding of include/b200sdr.h -- names, argument meaning and status codes mirror the C ABI
one to one (which in turn mirrors USBH_StatusTypeDef and the class callback shape of the
reference).  Nothing here computes: every method is one call into libb200sdr.so.
"""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200sdr.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "b200sdr.h")

OK, BUSY, FAIL, NOT_SUPPORTED, UNRECOVERED_ERROR = 0, 1, 2, 3, 4
CHAIN_SPECTRUM, CHAIN_WBFM, CHAIN_AM, CHAIN_COUNTER = 1, 2, 4, 8
WINDOW_RECT, WINDOW_HANN, WINDOW_BLACKMAN = 0, 1, 2
AVG_MEAN, AVG_EMA = 0, 1
SYNTH_COUNTER, SYNTH_MULTITONE, SYNTH_WBFM, SYNTH_AM = 0, 1, 2, 3
FIR_ENGINE_FP32, FIR_ENGINE_TENSOR = 0, 1

_STATUS = {0: "OK", 1: "BUSY", 2: "FAIL", 3: "NOT_SUPPORTED", 4: "UNRECOVERED_ERROR"}


class B200SdrError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: B200SDR_{_STATUS.get(status, status)} {detail}".strip())


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("device", C.c_int32),
        ("chains", C.c_uint32),
        ("window", C.c_uint32),
        ("avg_mode", C.c_uint32),
        ("ema_beta", C.c_float),
        ("ring_slots", C.c_uint32),
        ("slot_bytes", C.c_uint32),
        ("audio_capacity", C.c_uint32),
        ("submit_bytes", C.c_uint32),
        ("fir_engine", C.c_uint32),
        ("reserved", C.c_uint32 * 5),
    ]


_lib = None


def declared_symbols():
    """Every function include/b200sdr.h declares with B200SDR_API (for the export check)."""
    inc = os.path.dirname(HEADER_PATH)
    text = open(HEADER_PATH).read() + open(os.path.join(inc, "b200sdr_frontend.h")).read()
    return sorted(set(re.findall(r"B200SDR_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)))


def load_library():
    """dlopen the in-tree library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200SdrError(FAIL, "load_library", f"{LIB_PATH} not built (run __graft_entry__.build())")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, f32p = C.c_void_p, C.c_void_p, C.c_void_p
    u32, u64, i32 = C.c_uint32, C.c_uint64, C.c_int32
    sig = {
        "b200sdr_default_config": (None, [C.POINTER(Config)]),
        "b200sdr_create": (i32, [C.POINTER(Config), C.POINTER(vp)]),
        "b200sdr_destroy": (i32, [vp]),
        "process_samples": (i32, [u8p, u32, vp]),
        "b200sdr_ring_acquire": (i32, [vp, C.POINTER(vp), C.POINTER(u32)]),
        "b200sdr_ring_commit": (i32, [vp, u32]),
        "b200sdr_sync": (i32, [vp]),
        "b200sdr_reset": (i32, [vp]),
        "b200sdr_get_spectrum": (i32, [vp, f32p, C.POINTER(u64)]),
        "b200sdr_get_audio": (i32, [vp, u32, f32p, u32, C.POINTER(u32)]),
        "b200sdr_get_counters": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "b200sdr_batch_spectrum_dev": (i32, [vp, u8p, u32, u64, f32p]),
        "b200sdr_batch_wbfm_dev": (i32, [vp, u8p, u32, u64, f32p, f32p]),
        "b200sdr_batch_am_dev": (i32, [vp, u8p, u32, u64, f32p]),
        "b200sdr_batch_host": (i32, [vp, u32, u8p, u32, u64, f32p, f32p, f32p]),
        "b200sdr_spectrum_frames": (u64, [u64]),
        "b200sdr_wbfm_disc_len": (u64, [u64]),
        "b200sdr_wbfm_audio_len": (u64, [u64]),
        "b200sdr_am_audio_len": (u64, [u64]),
        "b200sdr_stream_chunk_samples": (u32, [u32]),
        "b200sdr_convert_cf32": (i32, [vp, u8p, u32, u32, f32p]),
        "b200sdr_convert_cf32_dev": (i32, [vp, u8p, u64, u32, f32p]),
        "b200sdr_counter_check_dev": (i32, [vp, u8p, u32, u64, C.c_int32, C.POINTER(u64), C.POINTER(u64)]),
        "b200sdr_counter_check": (i32, [vp, u8p, u32, C.c_int32, C.POINTER(u64), C.POINTER(u64)]),
        "b200sdr_get_counter_check": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "b200sdr_get_taps": (i32, [vp, u32, f32p, u32, C.POINTER(u32)]),
        "b200sdr_get_window": (i32, [vp, u32, f32p]),
        "b200sdr_debug_last_block": (i32, [vp, u8p, u32, C.POINTER(u32)]),
        "b200sdr_debug_wbfm_tc_acc": (i32, [vp, u8p, u64, vp, vp, C.POINTER(C.c_int32)]),
        "b200sdr_synth_fill_dev": (i32, [vp, u8p, u32, u64, u32, u64]),
        "b200sdr_synth_fill_host": (i32, [u8p, u32, u64, u32, u64]),
        "b200sdr_dev_alloc": (i32, [vp, u64, C.POINTER(vp)]),
        "b200sdr_dev_free": (i32, [vp, vp]),
        "b200sdr_host_alloc_pinned": (i32, [vp, u64, C.POINTER(vp)]),
        "b200sdr_host_free_pinned": (i32, [vp, vp]),
        "b200sdr_copy_to_host": (i32, [vp, vp, vp, u64]),
        "b200sdr_copy_to_dev": (i32, [vp, vp, vp, u64]),
        "b200sdr_timer_start": (i32, [vp]),
        "b200sdr_timer_stop_ms": (i32, [vp, C.POINTER(C.c_float)]),
        "b200sdr_kernel_launches": (u64, [vp]),
        "b200sdr_last_error": (C.c_char_p, [vp]),
        "b200sdr_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def spectrum_frames(len_bytes):
    return int(load_library().b200sdr_spectrum_frames(len_bytes))


def wbfm_disc_len(len_bytes):
    return int(load_library().b200sdr_wbfm_disc_len(len_bytes))


def wbfm_audio_len(len_bytes):
    return int(load_library().b200sdr_wbfm_audio_len(len_bytes))


def am_audio_len(len_bytes):
    return int(load_library().b200sdr_am_audio_len(len_bytes))


def stream_chunk_samples(chain):
    return int(load_library().b200sdr_stream_chunk_samples(chain))


def wbfm_stream_audio_len(total_bytes):
    """WBFM audio samples the streaming path has produced after `total_bytes` accepted bytes (whole chunks only)."""
    ch = stream_chunk_samples(CHAIN_WBFM)
    return -(-((total_bytes // 2 // ch) * (ch // 10)) // 5)


def synth_fill_host(n_captures, len_each, kind, first_capture=0):
    """Synthetic captures generated on the host by the library's own generator (no GPU needed)."""
    buf = np.empty(n_captures * len_each, dtype=np.uint8)
    rc = load_library().b200sdr_synth_fill_host(buf.ctypes.data, n_captures, len_each, kind, first_capture)
    if rc != OK:
        raise B200SdrError(rc, "b200sdr_synth_fill_host")
    return buf


class RtlRate(C.Structure):
    _fields_ = [("rsamp_ratio", C.c_uint32), ("real_rsamp_ratio", C.c_uint32), ("real_rate", C.c_double)]


class E4kPll(C.Structure):
    _fields_ = [("fosc", C.c_uint32), ("intended_flo", C.c_uint32), ("flo", C.c_uint32), ("x", C.c_uint16),
                ("z", C.c_uint8), ("r", C.c_uint8), ("r_idx", C.c_uint8), ("threephase", C.c_uint8)]


def rtl_resampler(samp_rate, xtal_hz=28800000):
    """include/b200sdr_frontend.h: (status, rsamp_ratio, real_rsamp_ratio, real_rate)."""
    lib = load_library()
    lib.b200sdr_rtl_resampler.restype = C.c_int32
    lib.b200sdr_rtl_resampler.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(RtlRate)]
    out = RtlRate()
    rc = lib.b200sdr_rtl_resampler(samp_rate, xtal_hz, C.byref(out))
    return rc, out.rsamp_ratio, out.real_rsamp_ratio, out.real_rate


def rtl_fir_pack(coeff=None):
    lib = load_library()
    lib.b200sdr_rtl_default_fir.argtypes = [C.POINTER(C.c_int32)]
    lib.b200sdr_rtl_fir_pack.restype = C.c_int32
    lib.b200sdr_rtl_fir_pack.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_uint8)]
    c = (C.c_int32 * 16)()
    if coeff is None:
        lib.b200sdr_rtl_default_fir(c)
    else:
        c[:] = list(coeff)
    out = (C.c_uint8 * 20)()
    rc = lib.b200sdr_rtl_fir_pack(c, out)
    return rc, bytes(out), list(c)


class CtlXfer(C.Structure):
    _fields_ = [("bmRequestType", C.c_uint8), ("bRequest", C.c_uint8), ("wValue", C.c_uint16), ("wIndex", C.c_uint16),
                ("wLength", C.c_uint16), ("data", C.c_uint8 * 2), ("step", C.c_uint8), ("reserved", C.c_uint8)]


def rtl_init_sequence(samp_rate=240000, xtal_hz=28800000, fir=None, flags=1, capacity=256):
    """include/b200sdr_frontend.h: (status, n, raw 12-byte records as bytes)."""
    lib = load_library()
    lib.b200sdr_rtl_init_sequence.restype = C.c_int32
    lib.b200sdr_rtl_init_sequence.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32, C.POINTER(CtlXfer),
                                              C.c_uint32, C.POINTER(C.c_uint32)]
    c = None
    if fir is not None:
        c = (C.c_int32 * 16)(*fir)
    out = (CtlXfer * max(capacity, 1))()
    n = C.c_uint32(0)
    rc = lib.b200sdr_rtl_init_sequence(samp_rate, xtal_hz, c, flags, out if capacity else None, capacity, C.byref(n))
    return rc, n.value, bytes(out)[: 12 * min(n.value, capacity)]


def e4k_pll_params(fosc, intended_flo):
    lib = load_library()
    lib.b200sdr_e4k_pll_params.restype = C.c_uint32
    lib.b200sdr_e4k_pll_params.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(E4kPll)]
    p = E4kPll()
    flo = lib.b200sdr_e4k_pll_params(fosc, intended_flo, C.byref(p))
    return flo, (p.fosc, p.intended_flo, p.flo, p.x, p.z, p.r, p.r_idx, p.threephase)


def e4k_band(flo_hz):
    lib = load_library()
    lib.b200sdr_e4k_band.restype = C.c_int32
    lib.b200sdr_e4k_band.argtypes = [C.c_uint32]
    return lib.b200sdr_e4k_band(flo_hz)


def e4k_rf_filter(band, freq_hz):
    lib = load_library()
    lib.b200sdr_e4k_rf_filter.restype = C.c_int32
    lib.b200sdr_e4k_rf_filter.argtypes = [C.c_int32, C.c_uint32]
    return lib.b200sdr_e4k_rf_filter(band, freq_hz)


def e4k_if_bw_index(filt, bw_hz):
    """-> (register index, bandwidth in Hz that index selects)"""
    lib = load_library()
    lib.b200sdr_e4k_if_bw_index.restype = C.c_int32
    lib.b200sdr_e4k_if_bw_index.argtypes = [C.c_int32, C.c_uint32, C.POINTER(C.c_uint32)]
    hz = C.c_uint32(0)
    idx = lib.b200sdr_e4k_if_bw_index(filt, bw_hz, C.byref(hz))
    return idx, hz.value


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class B200Sdr:
    """One context = one GPU, like one USBH host handle = one dongle in the reference."""

    def __init__(self, device=0, chains=CHAIN_SPECTRUM | CHAIN_WBFM | CHAIN_AM, window=WINDOW_HANN,
                 avg_mode=AVG_MEAN, ema_beta=0.1, ring_slots=8, slot_bytes=262144, audio_capacity=1 << 20,
                 submit_bytes=0, fir_engine=FIR_ENGINE_FP32):
        self.lib = load_library()
        cfg = Config()
        self.lib.b200sdr_default_config(C.byref(cfg))
        cfg.device, cfg.chains, cfg.window, cfg.avg_mode = device, chains, window, avg_mode
        cfg.ema_beta, cfg.ring_slots, cfg.slot_bytes, cfg.audio_capacity = ema_beta, ring_slots, slot_bytes, audio_capacity
        cfg.submit_bytes = submit_bytes
        cfg.fir_engine = fir_engine
        self.cfg = cfg
        self.ctx = C.c_void_p()
        rc = self.lib.b200sdr_create(C.byref(cfg), C.byref(self.ctx))
        if rc != OK:
            self.ctx = C.c_void_p()
            raise B200SdrError(rc, "b200sdr_create", "(needs a CUDA sm_100 device; no CPU fallback)")

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.lib.b200sdr_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where, allow_busy=False):
        if rc == OK or (allow_busy and rc == BUSY):
            return rc
        raise B200SdrError(rc, where, (self.lib.b200sdr_last_error(self.ctx) or b"").decode())

    # -- streaming --------------------------------------------------------------------------
    def process_samples(self, iq, allow_busy=False):
        iq = _u8(iq)
        rc = self.lib.process_samples(iq.ctypes.data, iq.size, self.ctx)
        return self._check(rc, "process_samples", allow_busy)

    def process_samples_raw(self, iq):
        """Status code only (for error-behaviour tests)."""
        iq = _u8(iq)
        return self.lib.process_samples(iq.ctypes.data, iq.size, self.ctx)

    def ring_acquire(self):
        ptr, nbytes = C.c_void_p(), C.c_uint32()
        rc = self.lib.b200sdr_ring_acquire(self.ctx, C.byref(ptr), C.byref(nbytes))
        if rc == BUSY:
            return None
        self._check(rc, "b200sdr_ring_acquire")
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes.value,))

    def ring_commit(self, nbytes, allow_busy=False):
        return self._check(self.lib.b200sdr_ring_commit(self.ctx, nbytes), "b200sdr_ring_commit", allow_busy)

    def sync(self):
        self._check(self.lib.b200sdr_sync(self.ctx), "b200sdr_sync")

    def reset(self):
        self._check(self.lib.b200sdr_reset(self.ctx), "b200sdr_reset")

    def get_spectrum(self):
        out = np.empty(1024, dtype=np.float32)
        n = C.c_uint64()
        self._check(self.lib.b200sdr_get_spectrum(self.ctx, out.ctypes.data, C.byref(n)), "b200sdr_get_spectrum")
        return out, int(n.value)

    def get_audio(self, chain, capacity=1 << 20):
        out = np.empty(capacity, dtype=np.float32)
        n = C.c_uint32()
        self._check(self.lib.b200sdr_get_audio(self.ctx, chain, out.ctypes.data, capacity, C.byref(n)), "b200sdr_get_audio")
        return out[: n.value].copy()

    def counters(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.b200sdr_get_counters(self.ctx, C.byref(a), C.byref(b), C.byref(c)), "b200sdr_get_counters")
        return {"bytes_in": a.value, "blocks_in": b.value, "busy_returns": c.value}

    def debug_last_block(self, capacity):
        out = np.empty(capacity, dtype=np.uint8)
        n = C.c_uint32()
        self._check(self.lib.b200sdr_debug_last_block(self.ctx, out.ctypes.data, capacity, C.byref(n)), "b200sdr_debug_last_block")
        return out[: min(n.value, capacity)].copy(), n.value

    def debug_wbfm_tc_acc(self, iq):
        """TENSOR FIR engine over one capture: (raw accumulators of the first tile [128][112] int32, tap slices [3][80] int8, e)."""
        iq = _u8(iq)
        d = self.dev_alloc(iq.size)
        try:
            self.to_dev(d, iq)
            acc = np.zeros((128, 112), np.int32)
            q = np.zeros((3, 80), np.int8)
            e = C.c_int32(0)
            self._check(self.lib.b200sdr_debug_wbfm_tc_acc(self.ctx, d, iq.size, acc.ctypes.data, q.ctypes.data, C.byref(e)),
                        "b200sdr_debug_wbfm_tc_acc")
        finally:
            self.dev_free(d)
        return acc, q, e.value

    # -- device memory ----------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.b200sdr_dev_alloc(self.ctx, nbytes, C.byref(p)), "b200sdr_dev_alloc")
        return p.value

    def dev_free(self, ptr):
        self._check(self.lib.b200sdr_dev_free(self.ctx, ptr), "b200sdr_dev_free")

    def pinned_alloc(self, nbytes, dtype=np.uint8):
        p = C.c_void_p()
        self._check(self.lib.b200sdr_host_alloc_pinned(self.ctx, nbytes, C.byref(p)), "b200sdr_host_alloc_pinned")
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))
        return p.value, arr.view(dtype)

    def pinned_free(self, ptr):
        self._check(self.lib.b200sdr_host_free_pinned(self.ctx, ptr), "b200sdr_host_free_pinned")

    def to_host(self, dev_ptr, nbytes, dtype=np.uint8):
        out = np.empty(nbytes, dtype=np.uint8)
        self._check(self.lib.b200sdr_copy_to_host(self.ctx, out.ctypes.data, dev_ptr, nbytes), "b200sdr_copy_to_host")
        return out.view(dtype)

    def to_dev(self, dev_ptr, host_array):
        a = np.ascontiguousarray(host_array)
        self._check(self.lib.b200sdr_copy_to_dev(self.ctx, dev_ptr, a.ctypes.data, a.nbytes), "b200sdr_copy_to_dev")

    # -- batched ----------------------------------------------------------------------------
    def batch_spectrum_dev(self, iq_dev, n_captures, len_each, out_dev):
        self._check(self.lib.b200sdr_batch_spectrum_dev(self.ctx, iq_dev, n_captures, len_each, out_dev), "b200sdr_batch_spectrum_dev")

    def batch_wbfm_dev(self, iq_dev, n_captures, len_each, audio_dev, disc_dev=None):
        self._check(self.lib.b200sdr_batch_wbfm_dev(self.ctx, iq_dev, n_captures, len_each, audio_dev, disc_dev), "b200sdr_batch_wbfm_dev")

    def batch_am_dev(self, iq_dev, n_captures, len_each, audio_dev):
        self._check(self.lib.b200sdr_batch_am_dev(self.ctx, iq_dev, n_captures, len_each, audio_dev), "b200sdr_batch_am_dev")

    def batch_host(self, chains, iq_host, n_captures, len_each, spectrum=None, wbfm=None, am=None):
        """iq_host / outputs: numpy arrays (ideally pinned, see pinned_alloc)."""
        def ptr(a):
            return a.ctypes.data if a is not None else None
        self._check(self.lib.b200sdr_batch_host(self.ctx, chains, iq_host.ctypes.data, n_captures, len_each,
                                                ptr(spectrum), ptr(wbfm), ptr(am)), "b200sdr_batch_host")

    # convenience wrappers used by the parity tests: host arrays in, host arrays out, all through the C ABI
    def spectrum(self, iq, n_captures=1):
        iq = _u8(iq)
        len_each = iq.size // n_captures
        out = np.empty((n_captures, 1024), dtype=np.float32)
        self.batch_host(CHAIN_SPECTRUM, iq, n_captures, len_each, spectrum=out)
        return out

    def wbfm(self, iq, n_captures=1, want_disc=False):
        iq = _u8(iq)
        len_each = iq.size // n_captures
        if not want_disc:
            out = np.empty((n_captures, wbfm_audio_len(len_each)), dtype=np.float32)
            self.batch_host(CHAIN_WBFM, iq, n_captures, len_each, wbfm=out)
            return out
        d_iq = self.dev_alloc(iq.size)
        na, nd = wbfm_audio_len(len_each), wbfm_disc_len(len_each)
        d_a, d_d = self.dev_alloc(4 * na * n_captures), self.dev_alloc(4 * nd * n_captures)
        try:
            self.to_dev(d_iq, iq)
            self.batch_wbfm_dev(d_iq, n_captures, len_each, d_a, d_d)
            self.sync()
            audio = self.to_host(d_a, 4 * na * n_captures, np.float32).reshape(n_captures, na)
            disc = self.to_host(d_d, 4 * nd * n_captures, np.float32).reshape(n_captures, nd)
        finally:
            self.dev_free(d_iq), self.dev_free(d_a), self.dev_free(d_d)
        return audio, disc

    def counter_check(self, iq, expect_first=-1):
        """b200sdr_counter_check (one host block): (n_breaks, first_break or None)."""
        iq = _u8(iq)
        n, f = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.b200sdr_counter_check(self.ctx, iq.ctypes.data if iq.size else None, iq.size, expect_first,
                                                   C.byref(n), C.byref(f)), "b200sdr_counter_check")
        return n.value, (None if f.value == 2**64 - 1 else f.value)

    def get_counter_check(self):
        """streaming totals (CHAIN_COUNTER): (n_breaks, first_break or None)."""
        n, f = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.b200sdr_get_counter_check(self.ctx, C.byref(n), C.byref(f)), "b200sdr_get_counter_check")
        return n.value, (None if f.value == 2**64 - 1 else f.value)

    def counter_check_dev(self, iq_dev, n_captures, len_each, expect_first=-1):
        """b200sdr_counter_check_dev: (n_breaks[n_captures], first_break[n_captures]) as uint64 arrays."""
        n = np.zeros(n_captures, np.uint64)
        f = np.zeros(n_captures, np.uint64)
        self._check(self.lib.b200sdr_counter_check_dev(self.ctx, iq_dev, n_captures, len_each, expect_first,
                                                       n.ctypes.data_as(C.POINTER(C.c_uint64)), f.ctypes.data_as(C.POINTER(C.c_uint64))),
                    "b200sdr_counter_check_dev")
        return n, f

    def am(self, iq, n_captures=1):
        iq = _u8(iq)
        len_each = iq.size // n_captures
        out = np.empty((n_captures, am_audio_len(len_each)), dtype=np.float32)
        self.batch_host(CHAIN_AM, iq, n_captures, len_each, am=out)
        return out

    # -- K2 / constants ---------------------------------------------------------------------
    def convert_cf32(self, iq, window=WINDOW_RECT):
        iq = _u8(iq)
        out = np.empty(iq.size, dtype=np.float32)
        self._check(self.lib.b200sdr_convert_cf32(self.ctx, iq.ctypes.data, iq.size, window, out.ctypes.data), "b200sdr_convert_cf32")
        return out

    def get_taps(self, which):
        out = np.empty(256, dtype=np.float32)
        n = C.c_uint32()
        self._check(self.lib.b200sdr_get_taps(self.ctx, which, out.ctypes.data, 256, C.byref(n)), "b200sdr_get_taps")
        return out[: n.value].copy()

    def get_window(self, window):
        out = np.empty(1024, dtype=np.float32)
        self._check(self.lib.b200sdr_get_window(self.ctx, window, out.ctypes.data), "b200sdr_get_window")
        return out

    def render_spectrum(self, spectrum=None, db_min=0.0, db_max=100.0):
        """480 x 272 ARGB8888 bar plot; spectrum=None renders the current streaming spectrum."""
        fn = self.lib.b200sdr_render_spectrum
        fn.restype = C.c_int32
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        img = np.empty((272, 480), dtype=np.uint32)
        src = None
        if spectrum is not None:
            spectrum = np.ascontiguousarray(spectrum, dtype=np.float32)
            src = spectrum.ctypes.data
        self._check(fn(self.ctx, src, db_min, db_max, img.ctypes.data), "b200sdr_render_spectrum")
        return img

    def render_waterfall(self, spectra, db_min=0.0, db_max=100.0):
        """480 x 272 ARGB8888 spectrogram: image row r shows spectra[r] (n_rows x 1024 float32)."""
        fn = self.lib.b200sdr_render_waterfall
        fn.restype = C.c_int32
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        spectra = np.ascontiguousarray(spectra, dtype=np.float32).reshape(-1, 1024)
        img = np.empty((272, 480), dtype=np.uint32)
        src = spectra.ctypes.data if spectra.shape[0] else None
        self._check(fn(self.ctx, src, spectra.shape[0], db_min, db_max, img.ctypes.data), "b200sdr_render_waterfall")
        return img

    # -- split-capture exchange (one capture across the GPUs of a box) ----------------------------
    def exchange_create(self, world, rank):
        """-> the 64-byte IPC handle of this rank's mailbox (gather them, pass to exchange_connect)."""
        fn = self.lib.b200sdr_exchange_create
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        h = (C.c_uint8 * 64)()
        self._check(fn(self.ctx, world, rank, h), "b200sdr_exchange_create")
        return bytes(h)

    def exchange_connect(self, handles):
        fn = self.lib.b200sdr_exchange_connect
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.c_char_p]
        self._check(fn(self.ctx, b"".join(handles)), "b200sdr_exchange_connect")

    def exchange_connect_local(self, peers):
        fn = self.lib.b200sdr_exchange_connect_local
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]
        arr = (C.c_void_p * len(peers))(*[p.ctx for p in peers])
        self._check(fn(self.ctx, arr), "b200sdr_exchange_connect_local")

    def split_spectrum_dev(self, iq_slice_dev, len_slice, frames_total, spectrum_dev):
        fn = self.lib.b200sdr_split_spectrum_dev
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        self._check(fn(self.ctx, iq_slice_dev, len_slice, frames_total, spectrum_dev), "b200sdr_split_spectrum_dev")

    def exchange_wait(self):
        fn = self.lib.b200sdr_exchange_wait
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p]
        self._check(fn(self.ctx), "b200sdr_exchange_wait")

    def exchange_destroy(self):
        fn = self.lib.b200sdr_exchange_destroy
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p]
        self._check(fn(self.ctx), "b200sdr_exchange_destroy")

    def synth_fill_dev(self, iq_dev, n_captures, len_each, kind, first_capture=0):
        self._check(self.lib.b200sdr_synth_fill_dev(self.ctx, iq_dev, n_captures, len_each, kind, first_capture), "b200sdr_synth_fill_dev")

    # -- timing -----------------------------------------------------------------------------
    def timer_start(self):
        self._check(self.lib.b200sdr_timer_start(self.ctx), "b200sdr_timer_start")

    def timer_stop_ms(self):
        ms = C.c_float()
        self._check(self.lib.b200sdr_timer_stop_ms(self.ctx, C.byref(ms)), "b200sdr_timer_stop_ms")
        return float(ms.value)

    def kernel_launches(self):
        return int(self.lib.b200sdr_kernel_launches(self.ctx))
