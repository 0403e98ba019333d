#!/bin/bash
# build several compile-time variants ON the box and bench each (spectrum / wbfm chain numbers)
mkdir -p gpurun_out
: > gpurun_out/variants.txt
while IFS= read -r v; do
  [ -z "$v" ] && continue
  B200_NVCC_EXTRA="$v" python stm32f7-rtlsdr_b200/build.py --force > /dev/null 2>&1 || { echo "BUILD FAILED: $v" >> gpurun_out/variants.txt; continue; }
  timeout 600 python bench.py --steps 3 --warmup 3 --captures-per-gpu 96 --e2e-captures 4 --no-cpu-baseline > gpurun_out/bench_var.txt 2>&1
  python - "$v" <<'PY' >> gpurun_out/variants.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/bench_var.txt').read().strip().splitlines()[-1])
    c=d['chains']
    print(f"{sys.argv[1]:70s} value {d['value']:9.0f}  spec {c['spectrum']['MSps_per_gpu']:9.0f} ({c['spectrum']['hbm_frac']:.4f})  wbfm {c['wbfm']['MSps_per_gpu']:9.0f} ({c['wbfm']['hbm_frac']:.4f})")
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/bench_var.txt').read()[-500:])
PY
done < tools/variants.list
cat gpurun_out/variants.txt
