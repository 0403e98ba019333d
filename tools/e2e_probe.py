"""Times b200sdr_batch_host call by call (variance of the end-to-end path) and plain H2D copies."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("stm32f7-rtlsdr_b200")
CB = 48_000_000
E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sdr = pkg.B200Sdr(chains=pkg.CHAIN_SPECTRUM | pkg.CHAIN_WBFM)
hp, h_iq = sdr.pinned_alloc(E * CB)
na = pkg.wbfm_audio_len(CB)
sp, h_spec = sdr.pinned_alloc(E * 1024 * 4, np.float32)
fp, h_fm = sdr.pinned_alloc(E * na * 4, np.float32)
h_iq[:CB] = pkg.synth_fill_host(1, CB, pkg.SYNTH_WBFM, 0)
for c in range(1, E):
    h_iq[c * CB:(c + 1) * CB] = h_iq[:CB]
d = sdr.dev_alloc(E * CB)
for i in range(12):
    t0 = time.perf_counter()
    sdr.batch_host(pkg.CHAIN_SPECTRUM | pkg.CHAIN_WBFM, h_iq, E, CB, spectrum=h_spec, wbfm=h_fm)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    sdr.to_dev(d, h_iq)
    dt2 = time.perf_counter() - t1
    print(f"call {i:2d}: batch_host {dt*1e3:7.2f} ms = {E*CB/2/dt/1e6:8.0f} MS/s ({E*CB/dt/1e9:5.1f} GB/s)   plain H2D {dt2*1e3:7.2f} ms ({E*CB/dt2/1e9:5.1f} GB/s)")
