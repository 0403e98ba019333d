#!/bin/bash
# SASS opcode mix per kernel of the SHIPPED library (cuobjdump -sass stm32f7-rtlsdr_b200/libb200sdr.so): static instruction
# counts of the opcodes that identify the formulation (packed fp32x2, TMA, tcgen05 / TMEM), plus the total.
#   bash tools/sass_mix_lib.sh > profiles/r2_sass_mix.txt
cd "$(dirname "$0")/.." || exit 1
LIB=stm32f7-rtlsdr_b200/libb200sdr.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  $LIB  ($(stat -c %s $LIB) bytes)  cuobjdump -sass, static counts per kernel"
echo "# FFMA2/FADD2/FMUL2 = packed fp32x2; UBLKCP = cp.async.bulk (1-D TMA); UTMALDG = cp.async.bulk.tensor (tensor-map TMA);"
echo "# UTCIMMA = tcgen05.mma kind::i8; LDTM = tcgen05.ld; UTCBAR = tcgen05.commit; SYNCS = mbarrier ops; LDGSTS = cp.async"
cuobjdump -sass $LIB | awk '
/Function :/ { f=$3; next }
$1 ~ /^\/\*[0-9a-f]+\*\/$/ {
  op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); sub(/;$/,"",op); n[f]++; c[f","op]++; ops[op]=1 }
END {
  split("FFMA2 FADD2 FMUL2 FFMA FADD FMUL HFMA2 I2FP PRMT MUFU LDG LDS STS SHFL BAR UBLKCP UTMALDG LDGSTS SYNCS UTCIMMA UTCBAR LDTM", want, " ")
  for (f in n) {
    line=sprintf("%-44s total %5d :", f, n[f])
    for (i=1;i<=22;i++) { k=f","want[i]; if (c[k]>0) line=line sprintf(" %s %d", want[i], c[k]) }
    print line
  }
}' | sort
