#!/bin/bash
# A/B of the u8->f32 conversion forms (cplx2.cuh B200_CVT_FHADD): microbenchmarks, the two prebuilt
# libraries benched on the same box, then the GPU parity tests with form B in place.
mkdir -p gpurun_out
./build/ubench > gpurun_out/ubench.txt 2>&1; grep -E "FHADD|cvt|HFMA2|^FFMA2  |^PRMT |subnormal" gpurun_out/ubench.txt
VCAPS=${VCAPS:-128} bash tools/gpu_variants.sh
LIB=stm32f7-rtlsdr_b200/libb200sdr.so
cp $LIB build/variants/original.so
cp build/variants/2.so $LIB
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_cvtB.txt
cp build/variants/original.so $LIB
