#!/bin/bash
# SASS instruction count (and bytes) per de_psd_kernel instantiation of a libeegfe build: keep the hot kernels < 32 KB (L1.5 I-cache).
for f in $(cuobjdump -sass "$1" | grep "Function :" | grep 13de_psd_kernelI | awk '{print $3}'); do
  n=$(cuobjdump -sass -fun $f "$1" 2>/dev/null | grep -cE '^\s+/\*[0-9a-f]{4,5}\*/')
  echo "$(echo $f | sed -E 's/.*3Cfg(I[^b]*)Lb.*(Lb[01])EEvNS.*/\1 \2/') instrs=$n bytes=$((n*16))"
done
