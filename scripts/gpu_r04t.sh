#!/bin/bash
mkdir -p gpurun_out
T=r04t
for v in default u2mb5 u1mb5 u3mb5 u2mb6 u2mb5tw0 f1mb5 u2mb5 default; do
  if [ $v = default ]; then L=""; else L="giwaxsim_b200/_variants/libgiwaxs_b200_$v.so"; fi
  GIWAXS_B200_LIB=$L timeout 300 python scripts/time_fused.py 1e7 4096 256 3 > gpurun_out/${T}_$v.log 2>&1
  echo "$v: $(tail -1 gpurun_out/${T}_$v.log | cut -c1-70)"
done
