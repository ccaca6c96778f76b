#!/bin/bash
# Round 2, call L (1 GPU): TMA-brick detector variant vs the LDG kernel; e2e with background-populated results
mkdir -p gpurun_out
T=r04l
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${T}_pytest.log
for mode in plain generic; do
  timeout 300 python scripts/time_detector.py 2048 360 $mode affine,affine_tma,affine,affine_tma > gpurun_out/${T}_detector_$mode.log 2>&1
  cat gpurun_out/${T}_detector_$mode.log | cut -c1-200
done
timeout 300 ncu --set full --clock-control none -k regex:'detector_affine_brick_kernel|detector_affine_kernel' -s 4 -c 2 \
    -o gpurun_out/${T}_det -f python scripts/time_detector.py 2048 360 generic affine,affine_tma > gpurun_out/${T}_det_ncu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
tail -4 gpurun_out/${T}_trace_e2e.log
