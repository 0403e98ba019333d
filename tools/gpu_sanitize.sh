#!/bin/bash
# compute-sanitizer over the small parity tests (memcheck + racecheck + synccheck), then a mid bench incl. AM
mkdir -p gpurun_out
SEL="config0 or golden_fixture or convert_bit_exact_all or (wbfm_batches and 2416) or (am_batches and 4816) or reset_starts or ring_acquire"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -n 6 gpurun_out/sanitizer_$tool.txt
done
timeout 900 python bench.py --steps 3 --warmup 3 --captures-per-gpu 128 --e2e-captures 32 --no-cpu-baseline > gpurun_out/bench_small.txt 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_small.txt').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']))
for k,v in d['chains'].items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
