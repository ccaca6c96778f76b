#!/bin/bash
# Round 2, call F (1 GPU): host-side profile of the two drivers + post-hoc transform tests
mkdir -p gpurun_out
T=r04f
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
timeout 600 python scripts/trace_e2e.py > gpurun_out/${T}_cprofile.log 2>&1
grep -E "wall|cumulative" -A28 gpurun_out/${T}_cprofile.log | head -90
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
tail -8 gpurun_out/${T}_trace_e2e.log
