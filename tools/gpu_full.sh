#!/bin/bash
# full default run as the driver does it: GPU tests, smoke, reference arm, default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -n 3 gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_default.txt 2> gpurun_out/bench_default.err; echo "bench wall seconds: $(( $(date +%s) - T0 ))"; tail -n 3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.txt').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'])
for k,v in d['chains'].items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='note'})
print('cpu', d['cpu_baseline'])
PY
