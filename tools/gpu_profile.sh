#!/bin/bash
# round-1 final profiles: launch lists (profile-size and default bench command) + ncu --set full of the three kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --captures-per-gpu 32 --e2e-captures 4 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_r1_c32.csv $B > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 512 -c 60 --csv --log-file gpurun_out/launches_r1_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_default.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_spectrum$' -s 3 -c 1 -f -o gpurun_out/prof_spectrum_r1 $B > gpurun_out/ncu_spec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_wbfm$' -s 3 -c 1 -f -o gpurun_out/prof_wbfm_r1 $B > gpurun_out/ncu_wbfm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_am_front$' -s 1 -c 1 -f -o gpurun_out/prof_am_r1 $B > gpurun_out/ncu_am.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_counter_check$' -s 1 -c 1 -f -o gpurun_out/prof_counter_r1 $B > gpurun_out/ncu_counter.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err
tail -c 400 gpurun_out/bench_default.txt
ls -la gpurun_out | grep -E "ncu-rep|launches"
