"""Worst UNRELAXED relative error per bin of the spectrum chain when only a few frames are averaged
(tests/test_gpu_parity.py uses 1e-5 * P + 4e-7 * sqrt(P * Pmax) below 200 frames).
Prints one line per case: frames, worst |err| / P over all bins, the bin, that bin's power relative to the strongest
bin, and the worst excess over the relaxed bound.  Run on the GPU box: python tools/few_frames_error.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_api import SYNTH_MULTITONE, SYNTH_WBFM, Golden  # noqa: E402

pkg = importlib.import_module("stm32f7-rtlsdr_b200")
g = Golden()
vec = np.load(os.path.join(ROOT, "tests", "golden", "golden_vectors.npz"))
with pkg.B200Sdr(chains=pkg.CHAIN_SPECTRUM) as sdr:
    cases = [("golden fixture spec_hann_mean", pkg.synth_fill_host(1, int(vec["spec_len"]), SYNTH_MULTITONE, int(vec["spec_seed"])), vec["spec_hann_mean"])]
    for n_frames, kind, seed in ((1, SYNTH_MULTITONE, 40), (4, SYNTH_WBFM, 77), (15, SYNTH_MULTITONE, 41), (63, SYNTH_MULTITONE, 42),
                                 (255, SYNTH_MULTITONE, 0), (1023, SYNTH_MULTITONE, 43), (46874, SYNTH_MULTITONE, 1000)):
        nbytes = 2 * (1024 + 512 * (n_frames - 1))
        nbytes += (-nbytes) % 16
        iq = g.synth(1, nbytes, kind, seed)
        cases.append((f"{n_frames} frames, synth kind {kind} seed {seed}", iq, g.spectrum(iq)[0]))
    for name, iq, gold in cases:
        out = sdr.spectrum(iq)[0].astype(np.float64)
        err = np.abs(out - gold)
        rel = err / gold
        k = int(rel.argmax())
        bound = 1e-5 * gold + 4e-7 * np.sqrt(gold * gold.max())
        frames = (iq.size // 2 - 1024) // 512 + 1
        print(f"{name:45s} frames {frames:6d}  worst rel err {rel.max():.3e} at bin {k:4d} (P_k / P_max = {gold[k] / gold.max():.2e})  "
              f"rel err of the strongest bin {rel[int(gold.argmax())]:.2e}  worst err / relaxed bound {np.max(err / bound):.3f}")
