#!/bin/bash
# SASS instruction mix (top opcodes + total) of the three hot kernels for every build/variants/<n>.so
cd "$(dirname "$0")/../build/variants" || exit 1
for so in $(ls [0-9]*.so | sort -V); do
  n=${so%.so}; cuobjdump -sass $so > $n.sass
  echo "== $n: $(cat $n.flags)"
  for k in k_spectrumILb0 k_wbfm k_am_front; do
    awk -v k=$k '/Function :/{f=index($0,k)>0} f' $n.sass | grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${TOP:-9} | tr '\n' ' '
    awk -v k=$k '/Function :/{f=index($0,k)>0} f' $n.sass | grep -cE "^\s+/\*[0-9a-f]{4}\*/"
  done
done
