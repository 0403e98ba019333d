"""The C-ABI library: loads, exports every symbol include/b200sdr.h declares, host-only entry
points agree with the oracle, status codes mirror USBH_StatusTypeDef.  No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAVE_GPU
from oracle_api import SYNTH_AM, SYNTH_COUNTER, SYNTH_MULTITONE, SYNTH_WBFM, Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(sdr_lib):
    names = sdr_lib.declared_symbols()
    assert "process_samples" in names and "b200sdr_create" in names and len(names) >= 35
    lib = C.CDLL(sdr_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    sdr_lib.load_library()  # binds every signature; raises on a missing export


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "b200sdr.h"\n#include "b200sdr_synth.h"\nint main(void){b200sdr_config c; (void)c; return 0;}\n')
    import subprocess
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_status_codes_are_usbh_status_values(sdr_lib):
    # USBH_StatusTypeDef: OK 0, BUSY 1, FAIL 2, NOT_SUPPORTED 3, UNRECOVERED_ERROR 4 (usbh_def.h:301-310)
    hdr = open(os.path.join(ROOT, "include", "b200sdr.h")).read()
    vals = dict(re.findall(r"#define B200SDR_(OK|BUSY|FAIL|NOT_SUPPORTED|UNRECOVERED_ERROR)\s+(\d+)", hdr))
    assert vals == {"OK": "0", "BUSY": "1", "FAIL": "2", "NOT_SUPPORTED": "3", "UNRECOVERED_ERROR": "4"}
    assert (sdr_lib.OK, sdr_lib.BUSY, sdr_lib.FAIL, sdr_lib.NOT_SUPPORTED) == (0, 1, 2, 3)


@pytest.mark.parametrize("nbytes", [0, 4, 2044, 2048, 2052, 3072, 262144, 48000000, 1208 * 2])
def test_output_lengths_match_oracle(sdr_lib, nbytes):
    g = Golden()
    n = nbytes // 2
    assert sdr_lib.spectrum_frames(nbytes) == (0 if n < 1024 else (n - 1024) // 512 + 1)
    assert sdr_lib.wbfm_disc_len(nbytes) == g.lib.gold_wbfm_disc_len(n)
    assert sdr_lib.wbfm_audio_len(nbytes) == g.lib.gold_wbfm_audio_len(n)
    assert sdr_lib.am_audio_len(nbytes) == g.lib.gold_am_audio_len(n)


def test_baseline_sizes(sdr_lib):
    assert sdr_lib.spectrum_frames(48_000_000) == 46874   # 10 s capture
    assert sdr_lib.spectrum_frames(262144) == 255         # one 256 KiB block
    assert sdr_lib.wbfm_audio_len(48_000_000) == 480000   # 48 kHz x 10 s
    assert sdr_lib.am_audio_len(48_000_000) == 80000      # 8 kHz x 10 s


@pytest.mark.parametrize("kind", [SYNTH_COUNTER, SYNTH_MULTITONE, SYNTH_WBFM, SYNTH_AM])
def test_host_generator_is_bit_identical_to_oracle(sdr_lib, kind):
    a = sdr_lib.synth_fill_host(3, 4096, kind, first_capture=11)
    b = Golden().synth(3, 4096, kind, first_capture=11)
    assert np.array_equal(a, b)


def test_default_config(sdr_lib):
    lib = sdr_lib.load_library()
    cfg = sdr_lib.Config()
    lib.b200sdr_default_config(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(sdr_lib.Config)
    assert cfg.slot_bytes == 16 * 32 * 512  # DEFAULT_BUF_LENGTH, usbh_rtlsdr.h:277-278
    assert cfg.slot_bytes % 4 == 0 and cfg.ring_slots >= 2


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(sdr_lib):
    # without a CUDA device the product refuses to create a context: there is no CPU path
    with pytest.raises(sdr_lib.B200SdrError) as ei:
        sdr_lib.B200Sdr()
    assert ei.value.status == sdr_lib.FAIL


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "stm32f7-rtlsdr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libgolden" not in text and "oracle_api" not in text and "/oracle/" not in text, f
                assert not re.search(r'#include\s+"[^"]*oracle', text), f


def test_shard_ranges_cover_batch(sdr_lib):
    import importlib
    sh = importlib.import_module("stm32f7-rtlsdr_b200.sharding")
    for n in (0, 1, 7, 8, 511, 512, 4096):
        for w in (1, 2, 3, 4, 8):
            r = sh.all_shards(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert sh.all_shards(4096, 8)[3] == (1536, 2048)  # 512 captures per GPU on the 8-GPU box


def test_example_host_driver_builds_and_refuses_without_gpu(sdr_lib, tmp_path):
    """examples/host_driver.c is the C host driver of INTEGRATION.md: it must compile as plain C
    against the header and link against the library (run for real by the GPU suite)."""
    import subprocess
    exe = tmp_path / "host_driver"
    libdir = os.path.dirname(sdr_lib.LIB_PATH)
    subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "host_driver.c"),
                    "-L", libdir, "-lb200sdr", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    if not HAVE_GPU:
        res = subprocess.run([str(exe), "--synthetic", "wbfm", "100000"], capture_output=True, text=True, cwd=tmp_path)
        assert res.returncode == 4 and "CUDA sm_100 device is required" in res.stderr
