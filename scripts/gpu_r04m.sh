#!/bin/bash
mkdir -p gpurun_out
T=r04m
timeout 300 compute-sanitizer --tool memcheck python scripts/time_detector.py 1024 24 generic affine_tma > gpurun_out/${T}_sanitizer.log 2>&1
grep -E "Invalid|Error|ERROR|at 0x|by thread|detector_affine_brick|Illegal|illegal" gpurun_out/${T}_sanitizer.log | head -30
timeout 600 python -m pytest tests/test_gpu_properties.py -q -k "brick or two_engines" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${T}_pytest.log
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
tail -6 gpurun_out/${T}_trace_e2e.log
