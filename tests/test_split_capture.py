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


def _run_fused_exchange(sh, sdr_lib, g, ctxs, iq):
    """every 'rank' = one context: slice -> fused finalize + peer-memory all-reduce -> 1024 floats each"""
    world = len(ctxs)
    frames = (iq.size // 2 - 1024) // 512 + 1
    bufs, outs = [], []
    for r, s in enumerate(ctxs):  # allocations first: cudaMalloc may wait for kernels already spinning
        b0, b1, f0, f1 = sh.split_capture_bytes(iq.size, r, world)
        d_in = s.dev_alloc(max(b1 - b0, 16))
        d_out = s.dev_alloc(4096)
        if b1 > b0:
            s.lib.b200sdr_copy_to_dev(s.ctx, d_in, iq[b0:b1].ctypes.data, b1 - b0)
        bufs.append((d_in, d_out, b1 - b0))
    for s, (d_in, d_out, n) in zip(ctxs, bufs):
        s.split_spectrum_dev(d_in, n, frames, d_out)          # asynchronous: rank r+1 is launched while r waits
    bufs = [(a, b) for a, b, _ in bufs]
    for s, (d_in, d_out) in zip(ctxs, bufs):
        s.exchange_wait()
        outs.append(s.to_host(d_out, 4096, np.float32).copy())
        s.dev_free(d_in)
        s.dev_free(d_out)
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4, 7])
def test_fused_finalize_exchange_on_one_device(sh, sdr_lib, world):
    """The fused finalize + all-reduce kernel (K6) with the ranks as contexts of ONE device: the
    peer mailboxes are then plain device pointers, the protocol (push, flags, bounded wait, rank-order
    sum, parity double-buffering over repeated calls) is the one used across NVLink."""
    g = Golden()
    ctxs = [sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) for _ in range(world)]
    try:
        for r, s in enumerate(ctxs):
            s.exchange_create(world, r)
        for s in ctxs:
            s.exchange_connect_local(ctxs)
            # grow the per-context workspace now: on ONE device a cudaFree inside a later call would wait
            # for the other "ranks'" spinning kernels (separate processes / devices do not have this coupling)
            s.spectrum(np.zeros(8 * 262144, np.uint8))
        for rnd, nbytes in enumerate((4 * 262144, 262144 + 4096, 2048 + 1024 * 3, 8 * 262144)):
            iq = g.synth(1, nbytes, SYNTH_MULTITONE, 30 + rnd)
            whole, frames = g.spectrum(iq)
            outs = _run_fused_exchange(sh, sdr_lib, g, ctxs, iq)
            for o in outs[1:]:
                assert np.array_equal(o, outs[0])             # bitwise the same on every rank
            err = np.abs(outs[0].astype(np.float64) - whole)
            bound = 1e-5 * whole + (4e-7 * np.sqrt(whole * whole.max()) if frames < 200 else 0.0)
            assert np.all(err <= bound), (rnd, float(np.max(err / whole)))
    finally:
        for s in ctxs:
            s.close()


@pytest.mark.gpu
def test_exchange_error_behaviour(sdr_lib):
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM) as s:
        d = s.dev_alloc(4096)
        for call in (lambda: s.split_spectrum_dev(d, 4096, 7, d), lambda: s.exchange_connect([b"x" * 64]),
                     lambda: s.exchange_wait()):
            with pytest.raises(sdr_lib.B200SdrError) as ei:   # nothing created / connected yet
                call()
            assert ei.value.status == sdr_lib.FAIL
        for world, rank in ((0, 0), (17, 0), (4, 4)):
            with pytest.raises(sdr_lib.B200SdrError) as ei:
                s.exchange_create(world, rank)
            assert ei.value.status == sdr_lib.NOT_SUPPORTED
        s.exchange_create(1, 0)
        with pytest.raises(sdr_lib.B200SdrError):             # already created
            s.exchange_create(1, 0)
        s.exchange_connect([b""])
        for args in ((d, 4098, 7, d), (d, 4096, 0, d)):       # length not a multiple of 4 / no frames in total
            with pytest.raises(sdr_lib.B200SdrError) as ei:
                s.split_spectrum_dev(*args)
            assert ei.value.status == sdr_lib.NOT_SUPPORTED
        s.exchange_destroy()
        s.exchange_create(1, 0)                               # can be set up again after a destroy
        s.exchange_connect([b""])
        s.split_spectrum_dev(d, 0, 7, d)                      # a rank without data contributes zeros
        s.exchange_wait()
        assert not s.to_host(d, 4096, np.float32).any()
        s.dev_free(d)
    with sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM, avg_mode=1, ema_beta=0.1) as s:
        s.exchange_create(1, 0)
        s.exchange_connect([b""])
        d = s.dev_alloc(4096)
        with pytest.raises(sdr_lib.B200SdrError) as ei:       # EMA needs the frames in order: not splittable
            s.split_spectrum_dev(d, 4096, 7, d)
        assert ei.value.status == sdr_lib.NOT_SUPPORTED
        s.dev_free(d)


@pytest.mark.gpu
def test_fused_exchange_times_out_instead_of_hanging(sdr_lib):
    """A peer that never arrives: the kernel's wait is bounded, exchange_wait reports FAIL."""
    import time
    a, b = sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM), sdr_lib.B200Sdr(chains=sdr_lib.CHAIN_SPECTRUM)
    try:
        a.exchange_create(2, 0)
        b.exchange_create(2, 1)
        a.exchange_connect_local([a, b])
        b.exchange_connect_local([a, b])
        d_in, d_out = a.dev_alloc(4096), a.dev_alloc(4096)
        t0 = time.time()
        a.split_spectrum_dev(d_in, 4096, 7, d_out)             # rank 1 never calls
        with pytest.raises(sdr_lib.B200SdrError) as ei:
            a.exchange_wait()
        assert ei.value.status == sdr_lib.FAIL and 2.0 < time.time() - t0 < 30.0
        a.dev_free(d_in)
        a.dev_free(d_out)
    finally:
        a.close()
        b.close()
