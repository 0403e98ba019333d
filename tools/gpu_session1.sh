#!/bin/bash
# first GPU session: microbenchmarks, smoke, parity tests, a short bench, launch list, one ncu capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; free -g >> gpurun_out/gpu_info.txt
timeout 120 ./build/ubench > gpurun_out/ubench.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --captures-per-gpu 128 --e2e-captures 16 > gpurun_out/bench_small.txt 2>&1
tail -5 gpurun_out/smoke.txt gpurun_out/pytest_gpu.txt gpurun_out/ubench.txt
tail -c 3000 gpurun_out/bench_small.txt
