#!/bin/bash
mkdir -p gpurun_out
T=r04n
timeout 300 python scripts/time_detector.py 2048 360 generic affine,affine_tma,affine,affine_tma > gpurun_out/${T}_detector_generic.log 2>&1
cut -c1-160 gpurun_out/${T}_detector_generic.log | grep -v "^ "
timeout 300 python scripts/time_detector.py 2048 360 plain affine,affine_tma,affine,affine_tma > gpurun_out/${T}_detector_plain.log 2>&1
cut -c1-160 gpurun_out/${T}_detector_plain.log | grep -v "^ "
timeout 600 python -m pytest tests/test_gpu_properties.py -q -k "brick" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest.log
timeout 300 compute-sanitizer --tool memcheck python scripts/time_detector.py 2048 8 generic affine_tma > gpurun_out/${T}_sanitizer.log 2>&1
grep -E "Invalid|Error|ERROR SUMMARY|at 0x|by thread|Illegal|illegal|Misaligned" gpurun_out/${T}_sanitizer.log | head -12
