#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/ubench > gpurun_out/ubench2.txt 2>&1
B="python bench.py --steps 2 --warmup 3 --captures-per-gpu 32 --e2e-captures 4 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spectrum<' -s 3 -c 1 -f -o gpurun_out/prof_spectrum_r1 $B > gpurun_out/ncu_spec.log 2>&1
cat gpurun_out/ubench2.txt
