#!/bin/bash
# generic A/B: bench every prebuilt variant (tools/build_variants.sh) at two batch sizes on the same box
mkdir -p gpurun_out
for caps in ${ABCAPS:-512 96}; do
  echo "== captures per GPU: $caps" | tee -a gpurun_out/ab.txt
  VCAPS=$caps bash tools/gpu_variants.sh | tee -a gpurun_out/ab.txt
done
