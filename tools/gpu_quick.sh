#!/bin/bash
# quick iteration: GPU tests (fail fast) + mid-size bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1
tail -n 6 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 3 --warmup 3 --captures-per-gpu ${CAPS:-128} --e2e-captures ${E2E:-32} --config4-waves ${WAVES:-2} --no-cpu-baseline > gpurun_out/bench_small.txt 2> gpurun_out/bench_small.err
tail -n 5 gpurun_out/bench_small.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_small.txt').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks'])
    for k,v in d['chains'].items(): print(k, {a: round(b,4) for a,b in v.items() if isinstance(b, float)})
    print('parity', d.get('parity')); print('config4', d.get('config4_single_gpu')); print('e2e', d['e2e']); print('ingest', d.get('ingest'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_small.txt').read()[-3000:])
PY
