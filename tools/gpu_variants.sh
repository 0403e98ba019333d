#!/bin/bash
# bench the prebuilt compile-time variants (tools/build_variants.sh) one after the other on the box:
# each build/variants/<n>.so is put in the library's place, bench.py runs, the original is restored.
mkdir -p gpurun_out
: > gpurun_out/variants.txt
LIB=stm32f7-rtlsdr_b200/libb200sdr.so
cp $LIB build/variants/original.so
for so in $(ls build/variants/[0-9]*.so | sort -V); do
  v=$(cat ${so%.so}.flags)
  cp $so $LIB
  timeout 600 python bench.py --steps 3 --warmup 3 --captures-per-gpu ${VCAPS:-96} --e2e-captures 4 --config4-waves 0 --parity-captures 4 --no-cpu-baseline > gpurun_out/bench_var.txt 2>&1
  python - "$v" <<'PY' >> gpurun_out/variants.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/bench_var.txt').read().strip().splitlines()[-1])
    c=d['chains']
    print(f"{sys.argv[1]:70s} value {d['value']:9.0f}  spec {c['spectrum']['MSps_per_gpu']:9.0f} ({c['spectrum']['hbm_frac']:.4f})  wbfm {c['wbfm']['MSps_per_gpu']:9.0f} ({c['wbfm']['hbm_frac']:.4f})  am {c.get('am',{}).get('MSps_per_gpu',0):9.0f} ({c.get('am',{}).get('hbm_frac',0):.4f})")
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/bench_var.txt').read()[-500:])
PY
done
cp build/variants/original.so $LIB
cat gpurun_out/variants.txt
