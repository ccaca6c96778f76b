#!/bin/bash
# Round 2, call J (1 GPU): final build - parity suite, bench, launch list, full captures of the hot kernels
mkdir -p gpurun_out
T=r04j
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -c 500 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_bench_under_ncu.json 2> gpurun_out/${T}_bench_under_ncu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'slice_rows_fused|slice_cols_tma' -s 2 -c 2 \
    -o gpurun_out/${T}_fused -f python scripts/time_fused.py 1e7 4096 64 1 > gpurun_out/${T}_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'detector_affine_kernel|voxel_finalize_kernel' -s 2 -c 2 \
    -o gpurun_out/${T}_stageb -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_stageb.log 2>&1
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
tail -6 gpurun_out/${T}_trace_e2e.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_compare.py -x -q -k "single_slice or 4189 or 65535 or shift_peak or polar" > gpurun_out/${T}_memcheck.log 2>&1
tail -4 gpurun_out/${T}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_fullsize.py -x -q -k "single_slice" > gpurun_out/${T}_racecheck.log 2>&1
tail -4 gpurun_out/${T}_racecheck.log
ls -la gpurun_out | grep ${T}
