#!/bin/bash
# Round 2, call O (1 GPU): TMA-fed column kernel at N = 1024 and 2048; full suite; PCIe probe
mkdir -p gpurun_out
T=r04o
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
for N in 1024 2048 4096; do
  A=$((N*N/2))
  timeout 300 python scripts/time_fused.py $A $N 128 3 > gpurun_out/${T}_fused_tma_$N.log 2>&1
  GIWAXS_B200_NO_TMA=1 timeout 300 python scripts/time_fused.py $A $N 128 3 > gpurun_out/${T}_fused_ldg_$N.log 2>&1
  echo "N=$N tma: $(tail -1 gpurun_out/${T}_fused_tma_$N.log | cut -c1-90)"
  echo "N=$N ldg: $(tail -1 gpurun_out/${T}_fused_ldg_$N.log | cut -c1-90)"
done
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fullsize.py -x -q -k "config3 or config4 or single_slice" > gpurun_out/${T}_memcheck.log 2>&1
tail -3 gpurun_out/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_fullsize.py -x -q -k "config3_graphite_crystal" > gpurun_out/${T}_racecheck.log 2>&1
tail -3 gpurun_out/${T}_racecheck.log
timeout 120 python scripts/pcie_bandwidth.py > gpurun_out/${T}_pcie.log 2>&1; cat gpurun_out/${T}_pcie.log
timeout 300 python scripts/time_config1.py > gpurun_out/${T}_config1.log 2>&1; tail -3 gpurun_out/${T}_config1.log
