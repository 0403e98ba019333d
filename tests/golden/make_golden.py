"""Regenerates tests/golden/golden_vectors.npz from the float64 golden model (oracle/golden.c).

    python tests/golden/make_golden.py

The vectors pin the golden model itself (tests/test_oracle_golden.py re-derives them with
numpy/scipy) and serve as fixtures for the GPU parity tests on the box, where neither
/root/reference nor scipy-derived ground truth is recomputed.  Inputs are produced by the
bit-reproducible generator of include/b200sdr_synth.h, so only seeds and outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_api import Golden, SYNTH_MULTITONE, SYNTH_WBFM, SYNTH_AM, SYNTH_COUNTER, WIN_HANN, WIN_BLACKMAN, AVG_EMA  # noqa: E402


def main():
    g = Golden()
    out = {}
    # spectrum: 8 KiB+ of multitone -> 15 frames ; Hann mean, Blackman mean, Hann EMA
    n_spec = 2 * (1024 + 512 * 14)
    iq = g.synth(1, n_spec, SYNTH_MULTITONE, first_capture=101)
    out["spec_seed"], out["spec_len"] = 101, n_spec
    out["spec_hann_mean"], _ = g.spectrum(iq)
    out["spec_blackman_mean"], _ = g.spectrum(iq, window=WIN_BLACKMAN)
    out["spec_hann_ema"], _ = g.spectrum(iq, window=WIN_HANN, avg_mode=AVG_EMA, beta=0.1)
    # first bytes of each generator kind: pins the synthetic-capture definition itself
    for kind, name in ((SYNTH_COUNTER, "counter"), (SYNTH_MULTITONE, "multitone"), (SYNTH_WBFM, "wbfm"), (SYNTH_AM, "am")):
        out[f"synth_{name}_head"] = g.synth(1, 256, kind, first_capture=7).copy()
    # WBFM: 48 000 bytes (a multiple of 16 and of 240)
    n_fm = 48000
    iq = g.synth(1, n_fm, SYNTH_WBFM, first_capture=202)
    audio, disc = g.wbfm(iq, want_disc=True)
    out["fm_seed"], out["fm_len"] = 202, n_fm
    out["fm_audio"], out["fm_disc"] = audio, disc
    # AM: 96 000 bytes (multiple of 16 and of 400)
    n_am = 96000
    iq = g.synth(1, n_am, SYNTH_AM, first_capture=303)
    out["am_seed"], out["am_len"] = 303, n_am
    out["am_audio"] = g.am(iq)
    for t in range(5):
        out[f"taps{t}"] = g.taps(t)
    out["window_hann"] = g.window(WIN_HANN)
    out["window_blackman"] = g.window(WIN_BLACKMAN)
    # exact conversion of all 256 byte values
    out["convert_all_bytes"] = g.convert(np.arange(256, dtype=np.uint8))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_vectors.npz"), **out)
    print("wrote golden_vectors.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
