#!/bin/bash
# 2-GPU check of the torchrun contract
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --captures-per-gpu 128 --e2e-captures 16 > gpurun_out/bench_n2.txt 2>&1
tail -c 2500 gpurun_out/bench_n2.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.txt 2>&1
tail -c 600 gpurun_out/bench_ref_n2.txt
