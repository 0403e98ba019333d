#!/bin/bash
# round-2 profiles in one gpurun call: launch lists (profile-size and default bench), ncu --set full of k_spectrum,
# k_wbfm_tc and k_wbfm, the default bench line, few-frames error table, ingest probes per chain
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --captures-per-gpu 32 --e2e-captures 4 --config4-waves 0 --parity-captures 4 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_c32.csv $B > gpurun_out/ncu_launch_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_spectrum$' -s 3 -c 1 -f -o gpurun_out/prof_spectrum_r2 $B > gpurun_out/ncu_spec_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_wbfm_tc$' -s 3 -c 1 -f -o gpurun_out/prof_wbfm_tc_r2 $B > gpurun_out/ncu_wbfm_tc_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_wbfm$' -s 2 -c 1 -f -o gpurun_out/prof_wbfm_r2 $B > gpurun_out/ncu_wbfm_r2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 80 --csv --log-file gpurun_out/launches_r2_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --config4-waves 0 > gpurun_out/ncu_launch_default_r2.log 2>&1
python tools/few_frames_error.py > gpurun_out/few_frames.txt 2>&1
for ch in 1 2 3 7 15; do build/ingest_bench 262144 2048 $ch; done > gpurun_out/ingest_chains.txt 2>&1
build/ingest_bench 1048576 1024 3 >> gpurun_out/ingest_chains.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_r2.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err
kill $SMI
tail -c 600 gpurun_out/bench_default.txt
ls -la gpurun_out | grep -E "r2"
