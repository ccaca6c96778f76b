#!/bin/bash
# Round 2, call B: integer-accumulator row kernel + TMA-fed column kernel: parity first, then timing.
mkdir -p gpurun_out
T=r04b
echo "== TMA kernel, isolated (N = 4096 single slice vs oracle)"
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -k "single_slice or fused_equals" > gpurun_out/${T}_tma_first.log 2>&1
RC=$?; echo "tma first rc=$RC"; tail -15 gpurun_out/${T}_tma_first.log
if [ $RC -ne 0 ]; then
  echo "== TMA path failed: rerun with GIWAXS_B200_NO_TMA=1"
  export GIWAXS_B200_NO_TMA=1
fi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${T}_bench_n1.err
GIWAXS_B200_NO_TMA=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_notma.json 2> gpurun_out/${T}_bench_notma.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'slice_(rows_fused|cols_tma|cols_fused)|voxel_finalize' -s 2 -c 2 \
    -o gpurun_out/${T}_fused -f python scripts/time_fused.py 1e7 4096 64 1 > gpurun_out/${T}_fused.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fullsize.py -x -q -k "single_slice" > gpurun_out/${T}_memcheck.log 2>&1
tail -5 gpurun_out/${T}_memcheck.log
ls -la gpurun_out | grep ${T}
