"""Do the spectrum and WBFM chains overlap usefully when launched on different streams?

Both are FP32-pipe bound at ~75 % pipe utilisation when run alone; co-resident CTAs of the other
kernel could fill the idle issue slots.  Two contexts on one device (= two independent stream sets)
run the two chains on the same resident captures; wall time of `steps` rounds (device-synchronised)
is compared with the same work issued into one context (serial on one stream).

    python tools/concurrency_probe.py [captures] [steps]
"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    pkg = importlib.import_module("stm32f7-rtlsdr_b200")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    L = 48_000_000
    a = pkg.B200Sdr(device=0, chains=pkg.CHAIN_SPECTRUM)
    b = pkg.B200Sdr(device=0, chains=pkg.CHAIN_WBFM)
    iq = torch.empty(B * L, dtype=torch.uint8, device="cuda")
    spec = torch.empty(B * 1024, dtype=torch.float32, device="cuda")
    audio = torch.empty(B * pkg.wbfm_audio_len(L), dtype=torch.float32, device="cuda")
    for c in range(B):
        a.synth_fill_dev(iq.data_ptr() + c * L, 1, L, pkg.SYNTH_MULTITONE if c % 2 == 0 else pkg.SYNTH_WBFM, first_capture=c)
    a.sync()

    def run(mode, chunk):
        """chunk = captures per launch (smaller chunks interleave the two kernels more finely)"""
        def once():
            for c0 in range(0, B, chunk):
                n = min(chunk, B - c0)
                p = iq.data_ptr() + c0 * L
                fm_ctx = a if mode == "serial" else b
                a.batch_spectrum_dev(p, n, L, spec.data_ptr() + c0 * 4096)
                fm_ctx.batch_wbfm_dev(p, n, L, audio.data_ptr() + c0 * 4 * pkg.wbfm_audio_len(L))
        for _ in range(2):
            once()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            once()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        print(f"{mode:10s} chunk {chunk:4d}: {dt * 1e3:8.3f} ms/step  {B * L / 2 / dt / 1e6:10.0f} MS/s", flush=True)
        return dt

    base = run("serial", B)
    for chunk in (B, 32, 8):
        dt = run("two_streams", chunk)
        print(f"    -> {base / dt:.3f}x of serial")
    a.close()
    b.close()


if __name__ == "__main__":
    main()
