#!/bin/bash
mkdir -p gpurun_out
T=r04s
for v in default f1mb3 f1mb5 f1mb6 u2 u2mb5 default; do
  if [ $v = default ]; then L=""; else L="giwaxsim_b200/_variants/libgiwaxs_b200_$v.so"; fi
  GIWAXS_B200_LIB=$L timeout 300 python scripts/time_fused.py 1e7 4096 256 3 > gpurun_out/${T}_$v.log 2>&1
  echo "$v: $(tail -1 gpurun_out/${T}_$v.log | cut -c1-70)"
done
