"""Optional exchange step: one capture split in time across ranks, bin sums
all-reduced.  CPU: range logic + a real 2-rank gloo all-reduce of oracle-computed slice spectra.
GPU: the slices run through the device kernel (ranks emulated one after the other on one GPU)."""
import importlib
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle_api import SYNTH_MULTITONE, Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sh():
    sys.path.insert(0, ROOT)
    return importlib.import_module("stm32f7-rtlsdr_b200.sharding")


def test_slices_cover_every_frame_once(sh):
    for nbytes in (2048, 2048 + 1024 * 6, 262144, 48_000_000, 2044):
        n = nbytes // 2
        frames = 0 if n < 1024 else (n - 1024) // 512 + 1
        for world in (1, 2, 3, 8):
            seen = 0
            for r in range(world):
                b0, b1, f0, f1 = sh.split_capture_bytes(nbytes, r, world)
                assert f0 == seen
                seen = f1
                if f1 > f0:
                    assert b0 == 1024 * f0 and b1 == 1024 * (f1 - 1) + 2048 and b1 <= nbytes
                    assert (b1 - b0 - 2048) // 1024 + 1 == f1 - f0   # the slice holds exactly its frames
            assert seen == frames


def test_weighted_slice_means_equal_whole_capture(sh):
    g = Golden()
    iq = g.synth(1, 262144 + 4096, SYNTH_MULTITONE, 12)
    whole, frames = g.spectrum(iq)
    acc = np.zeros(1024)
    for r in range(3):
        b0, b1, f0, f1 = sh.split_capture_bytes(iq.size, r, 3)
        part, fr = g.spectrum(iq[b0:b1])
        assert fr == f1 - f0
        acc += part * (f1 - f0) / frames
    assert np.max(np.abs(acc - whole) / whole) < 1e-12


WORKER = textwrap.dedent("""
    import importlib, os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    from oracle_api import Golden, SYNTH_MULTITONE
    sh = importlib.import_module("stm32f7-rtlsdr_b200.sharding")
    dist.init_process_group("gloo", init_method="env://")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = Golden()
    iq = g.synth(1, 262144, SYNTH_MULTITONE, 5)
    whole, frames = g.spectrum(iq)
    b0, b1, f0, f1 = sh.split_capture_bytes(iq.size, rank, world)
    part, fr = g.spectrum(iq[b0:b1])
    t = torch.from_numpy(part.copy())
    sh.allreduce_split_spectrum(t, f1 - f0, frames, dist)
    err = float(np.max(np.abs(t.numpy() - whole) / whole))
    assert err < 1e-12, err
    if rank == 0: print("SPLIT_OK", frames, f0, f1)
    dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_two_rank_allreduce_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29519", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    assert "SPLIT_OK 255 0 128" in outs[0][0]


@pytest.mark.gpu
def test_split_capture_on_device(sh, sdr_lib):
    import torch
    g = Golden()
    iq = g.synth(1, 4 * 262144, SYNTH_MULTITONE, 21)
    whole, frames = g.spectrum(iq)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        acc = torch.zeros(1024, dtype=torch.float32)
        for r in range(4):  # the ranks of a 4-GPU split, run one after the other on this GPU
            b0, b1, f0, f1 = sh.split_capture_bytes(iq.size, r, 4)
            part = torch.from_numpy(s.spectrum(iq[b0:b1])[0].copy())
            acc += sh.allreduce_split_spectrum(part, f1 - f0, frames, None)
    assert np.max(np.abs(acc.numpy().astype(np.float64) - whole) / whole) <= 1e-5
