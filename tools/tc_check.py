"""Tensor-core FIR engine (csrc/wbfm_tc.cuh) on the GPU box: (1) the raw u8 x s8 -> s32 accumulators of the first tile
against the integer FIR computed with numpy -- bit for bit; (2) discriminator / audio against oracle B; (3) both engines
timed on resident captures.  python tools/tc_check.py [captures] [seconds]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_api import SYNTH_WBFM, Golden, wrap_phase  # noqa: E402

pkg = importlib.import_module("stm32f7-rtlsdr_b200")
g = Golden()
n_cap = int(sys.argv[1]) if len(sys.argv) > 1 else 64
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0

with pkg.B200Sdr(chains=pkg.CHAIN_WBFM, fir_engine=pkg.FIR_ENGINE_TENSOR) as sdr:
    n = 160 * 125 * 3 + 1232
    iq = g.synth(1, 2 * n, SYNTH_WBFM, 11)
    try:
        acc, q, e = sdr.debug_wbfm_tc_acc(iq)
    except Exception as ex:  # a bounded wait expired: say which, and stop
        print("TC_CHECK debug run failed:", ex)
        sys.exit(1)
    u = iq.astype(np.int64)
    want = np.zeros((125, 112), np.int64)
    used = np.zeros(112, bool)
    for s in range(3):
        for c in (0, 1):
            x = np.concatenate([np.zeros(96, np.int64), u[c::2]])          # x[n < 0] = 0 (zero-filled history bytes)
            full = np.convolve(x, q[s].astype(np.int64))                   # full[96 + n] = sum_t q[t] x[n - t]
            for j in range(17):                                            # output m = 16 r - 1 + j at sample 10 m
                m = 16 * np.arange(125) - 1 + j
                want[:, 34 * s + 2 * j + c] = full[96 + 10 * m]
                used[34 * s + 2 * j + c] = True
    bad = np.argwhere((acc[:125].astype(np.int64) != want) & used[None, :])
    print(f"TC_CHECK accumulators: {'bit-exact' if bad.size == 0 else f'{len(bad)} of {125 * 102} differ'} (e = {e})")
    if bad.size:
        for r, col in bad[:12]:
            print(f"  row {r} col {col} (slice {col // 34} j {col % 34 // 2} comp {col % 2}): got {acc[r, col]} want {want[r, col]}")
        print("  rows with errors:", sorted(set(bad[:, 0]))[:40])
        print("  cols with errors:", sorted(set(bad[:, 1]))[:112])
    audio, disc = sdr.wbfm(iq, want_disc=True)
    ga, gd = g.wbfm(iq, want_disc=True)
    derr = np.abs(wrap_phase(disc[0] - gd))
    print(f"TC_CHECK parity: disc err first 8 {derr[:8].max():.2e}, rest {derr[8:].max():.2e} rad; audio err {np.abs(audio[0] - ga).max():.2e}")

len_each = int(seconds * 2_400_000) * 2
for engine, name in ((pkg.FIR_ENGINE_FP32, "fp32"), (pkg.FIR_ENGINE_TENSOR, "tensor")):
    with pkg.B200Sdr(chains=pkg.CHAIN_WBFM, fir_engine=engine) as sdr:
        d_iq = sdr.dev_alloc(n_cap * len_each)
        na = pkg.wbfm_audio_len(len_each)
        d_a = sdr.dev_alloc(4 * na * n_cap)
        try:
            sdr.lib.b200sdr_synth_fill_dev(sdr.ctx, d_iq, n_cap, len_each, SYNTH_WBFM, 0)
            for _ in range(2):
                sdr.batch_wbfm_dev(d_iq, n_cap, len_each, d_a)
            sdr.sync()
            ms = []
            for _ in range(4):
                sdr.lib.b200sdr_timer_start(sdr.ctx)
                sdr.batch_wbfm_dev(d_iq, n_cap, len_each, d_a)
                import ctypes as C
                t = C.c_float()
                sdr.lib.b200sdr_timer_stop_ms(sdr.ctx, C.byref(t))
                ms.append(t.value)
            sdr.sync()
            a = sdr.to_host(d_a, 4 * na * 2, np.float32)
            best = min(ms)
            rate = n_cap * len_each / 2 / (best * 1e-3)
            print(f"TC_CHECK engine {name:6s}: {best:8.3f} ms for {n_cap} x {seconds} s  = {rate / 1e12:.3f} T samples/s = "
                  f"{rate * 2.08 / 1e9:.0f} GB/s ({rate * 2.08 / 1e9 / 6545.9:.3f} of 6545.9)  all ms {[round(x, 3) for x in ms]}  checksum {float(np.abs(a).sum()):.6f}")
        finally:
            sdr.dev_free(d_iq)
            sdr.dev_free(d_a)
