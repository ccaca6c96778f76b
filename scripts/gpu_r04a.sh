#!/bin/bash
# Round 2, call A: parity suite + bench + launch list + full captures of the hot kernels (final round-1 kernels).
mkdir -p gpurun_out
T=r04a
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${T}_bench_n1.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -c 500 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_bench_under_ncu.json 2> gpurun_out/${T}_bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:'slice_(rows|cols)_fused' -s 2 -c 2 \
    -o gpurun_out/${T}_fused -f python scripts/time_fused.py 1e7 4096 64 1 > gpurun_out/${T}_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'detector_affine_kernel|voxel_finalize_kernel' -s 2 -c 2 \
    -o gpurun_out/${T}_stageb -f python bench.py --phis 64 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_stageb.log 2>&1
ls -la gpurun_out | tail -20
