#!/bin/bash
# k_spectrum warp-start stagger sweep: every build/variants/<n>.so x every value of B200SDR_SPEC_STAGGER (cycles)
mkdir -p gpurun_out
: > gpurun_out/stagger.txt
LIB=stm32f7-rtlsdr_b200/libb200sdr.so
cp $LIB build/variants/original.so
for so in $(ls build/variants/[0-9]*.so | sort -V); do
  v=$(cat ${so%.so}.flags)
  cp $so $LIB
  for S in ${STAGGERS:-0 160 320 640 1275}; do
    B200SDR_SPEC_STAGGER=$S timeout 600 python bench.py --steps 3 --warmup 3 --captures-per-gpu ${VCAPS:-128} --e2e-captures 4 --config4-waves 0 --parity-captures 4 --no-cpu-baseline > gpurun_out/bench_var.txt 2>&1
    python - "$v stagger=$S" <<'PY' >> gpurun_out/stagger.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/bench_var.txt').read().strip().splitlines()[-1])
    c=d['chains']
    print(f"{sys.argv[1]:60s} value {d['value']:9.0f}  spec {c['spectrum']['MSps_per_gpu']:9.0f} ({c['spectrum']['hbm_frac']:.4f})  wbfm {c['wbfm']['MSps_per_gpu']:9.0f}  parity {d['parity']['bitwise']}")
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/bench_var.txt').read()[-500:])
PY
  done
done
cp build/variants/original.so $LIB
cat gpurun_out/stagger.txt
