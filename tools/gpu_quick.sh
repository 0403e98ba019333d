#!/bin/bash
# quick iteration: GPU tests (fail fast) + mid-size bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1
tail -n 4 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 3 --warmup 3 --captures-per-gpu ${CAPS:-128} --e2e-captures ${E2E:-32} --no-cpu-baseline > gpurun_out/bench_small.txt 2>&1
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_small.txt').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
    for k,v in d['chains'].items(): print(k, {a: round(b,4) for a,b in v.items() if isinstance(b, float)})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_small.txt').read()[-3000:])
PY
