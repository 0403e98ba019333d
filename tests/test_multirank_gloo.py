"""world_size-2 run of the multi-GPU host logic on CPU (gloo): capture sharding, barrier and
max-over-ranks timing reduction exactly as bench.py does them.  No data-path collective exists."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import importlib, os, sys
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    sh = importlib.import_module("stm32f7-rtlsdr_b200.sharding")
    dist.init_process_group("gloo", init_method="env://")
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = sh.shard_range(4096, rank, world)
    owned = torch.zeros(4096, dtype=torch.int32); owned[lo:hi] = 1
    dist.all_reduce(owned)                       # test-only: every capture owned exactly once
    assert int(owned.min()) == 1 and int(owned.max()) == 1
    t = torch.tensor([10.0 + rank]); dist.barrier(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == 10.0 + world - 1          # bench.py: time = max over ranks
    units = torch.tensor([float(hi - lo)]); dist.all_reduce(units)
    assert int(units) == 4096
    if rank == 0: print("GLOO_OK", lo, hi)
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    assert "GLOO_OK 0 2048" in outs[0][0]
