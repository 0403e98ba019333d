#!/bin/bash
mkdir -p gpurun_out
nproc; cat /sys/devices/system/node/online 2>/dev/null; numactl -H 2>/dev/null | head -5
for numa in 0 1; do for e in 16 32 64; do
B200_BENCH_NUMA=$numa timeout 600 python bench.py --steps 3 --warmup 3 --captures-per-gpu 64 --e2e-captures $e --no-cpu-baseline > gpurun_out/bench_e2e.txt 2>&1
python - $numa $e <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_e2e.txt').read().strip().splitlines()[-1])
print('numa',sys.argv[1],'E',sys.argv[2],'e2e', round(d['e2e']['value']), 'GB/s', round(d['e2e']['value']*2/1e3,1))
PY
done; done
