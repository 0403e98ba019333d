"""One long capture split in time across the GPUs of a box (kernel K6).

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port 29541 tools/split_capture_p2p.py

Every rank runs k_spectrum over its own time range; the bin sums are combined
  (a) by the fused finalize + peer-memory all-reduce kernel (b200sdr_split_spectrum_dev), and
  (b) for comparison by k_spectrum_finalize + an NCCL all-reduce (sharding.allreduce_split_spectrum).
Checks: (a) is bitwise identical on every rank, (a) and (b) match the float64 golden spectrum of the
whole capture to 1e-5; prints the latency of both for a 10 s capture and for a 256 KiB block (where the
exchange itself dominates).  torch.distributed is plumbing here: it gathers the 64-byte IPC handles.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("stm32f7-rtlsdr_b200")
    sh = importlib.import_module("stm32f7-rtlsdr_b200.sharding")
    sdr = pkg.B200Sdr(device=local, chains=pkg.CHAIN_SPECTRUM)

    handles = [None] * world
    dist.all_gather_object(handles, sdr.exchange_create(world, rank))
    sdr.exchange_connect(handles)

    report = {"world": world}
    for label, nbytes, iters in (("capture_10s", 48_000_000, 30), ("block_256KiB", 262144, 200)):
        iq = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        sdr.synth_fill_dev(iq.data_ptr(), 1, nbytes, pkg.SYNTH_MULTITONE, first_capture=7)   # same capture on every rank
        sdr.sync()
        frames = (nbytes // 2 - 1024) // 512 + 1
        b0, b1, f0, f1 = sh.split_capture_bytes(nbytes, rank, world)
        out_fused = torch.zeros(1024, dtype=torch.float32, device="cuda")
        out_nccl = torch.zeros(1024, dtype=torch.float32, device="cuda")

        def fused():
            sdr.split_spectrum_dev(iq.data_ptr() + b0, b1 - b0, frames, out_fused.data_ptr())

        def nccl():
            if b1 > b0:
                sdr.batch_spectrum_dev(iq.data_ptr() + b0, 1, b1 - b0, out_nccl.data_ptr())
            else:
                out_nccl.zero_()
            sdr.sync()                                   # library stream -> torch stream
            sh.allreduce_split_spectrum(out_nccl, f1 - f0, frames, dist)
            torch.cuda.synchronize()                     # result complete (and out_nccl free for the next call)

        for _ in range(3):
            fused()
            sdr.exchange_wait()
            nccl()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            fused()
            sdr.exchange_wait()                          # result complete: both variants are timed call by call
        t_fused = (time.perf_counter() - t0) / iters
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            nccl()
        torch.cuda.synchronize()
        t_nccl = (time.perf_counter() - t0) / iters
        tt = torch.tensor([t_fused, t_nccl], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)

        gathered = [torch.zeros_like(out_fused) for _ in range(world)]
        dist.all_gather(gathered, out_fused)
        bitwise = all(torch.equal(gathered[0], t) for t in gathered)
        entry = {"bytes": nbytes, "frames": frames, "fused_us": float(tt[0]) * 1e6, "finalize_plus_nccl_us": float(tt[1]) * 1e6,
                 "bitwise_identical_on_all_ranks": bool(bitwise)}
        if rank == 0:
            from oracle_api import Golden, SYNTH_MULTITONE
            gold, gframes = Golden().spectrum(Golden().synth(1, nbytes, SYNTH_MULTITONE, 7))
            assert gframes == frames
            entry["fused_max_rel_err"] = float(np.max(np.abs(out_fused.cpu().numpy().astype(np.float64) - gold) / gold))
            entry["nccl_max_rel_err"] = float(np.max(np.abs(out_nccl.cpu().numpy().astype(np.float64) - gold) / gold))
            assert bitwise and entry["fused_max_rel_err"] <= 1e-5 and entry["nccl_max_rel_err"] <= 1e-5, entry
        report[label] = entry
    if rank == 0:
        print("SPLIT_P2P " + json.dumps(report))
    dist.barrier()
    sdr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
