#!/bin/bash
# ncu --set full of ONE launch of kernel $K (regex) for every build/variants/<n>.so -> gpurun_out/prof_ab_<n>.ncu-rep
mkdir -p gpurun_out
LIB=stm32f7-rtlsdr_b200/libb200sdr.so
cp $LIB build/variants/original.so
B="python bench.py --steps 2 --warmup 3 --captures-per-gpu ${NCAPS:-32} --e2e-captures 4 --config4-waves 0 --parity-captures 2 --no-cpu-baseline"
for so in $(ls build/variants/[0-9]*.so | sort -V); do
  n=$(basename ${so%.so})
  cp $so $LIB
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${K:-^k_spectrum$}" -s ${SKIP:-3} -c 1 -f -o gpurun_out/prof_ab_$n $B > gpurun_out/ncu_ab_$n.log 2>&1
  tail -n 2 gpurun_out/ncu_ab_$n.log
done
cp build/variants/original.so $LIB
ls -la gpurun_out | grep prof_ab
