#!/bin/bash
# Cross-compile every line of tools/variants.list into build/variants/<n>.so HERE (no GPU needed), so the
# GPU box only benches them (tools/gpu_variants.sh).  Lines are extra nvcc flags, e.g. -DB200_SPEC_MINB=4.
set -e
cd "$(dirname "$0")/.."
rm -rf build/variants && mkdir -p build/variants
n=0
while IFS= read -r v; do
  [ -z "$v" ] && continue
  n=$((n + 1))
  echo "$v" > build/variants/$n.flags
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC,-fvisibility=hidden,-O2 \
      -Xptxas -v -diag-suppress 550 $v -o build/variants/$n.so stm32f7-rtlsdr_b200/csrc/api.cu stm32f7-rtlsdr_b200/csrc/frontend.cpp \
      > build/variants/$n.log 2>&1 || echo "BUILD FAILED: $v" ) &
  [ $((n % 6)) -eq 0 ] && wait
done < tools/variants.list
wait
grep -l "spill stores" build/variants/*.log > /dev/null && grep -H -A1 "k_spectrumILb0\|k_wbfm" build/variants/*.log | grep -E "registers|bytes spill" | grep -v " 0 bytes spill stores" || true
ls build/variants/*.so | wc -l
