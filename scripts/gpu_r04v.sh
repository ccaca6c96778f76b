#!/bin/bash
# Round 2, call V (1 GPU): FINAL build - parity suite, bench (x2), reference arm, launch list, full captures
mkdir -p gpurun_out
T=r04v
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
for k in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1_$k.json 2> gpurun_out/${T}_bench_n1_$k.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1_$k.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['roofline']['kernel_ms_per_slice'].items()}, round(d['e2e']['value']), [c[0] for c in d['e2e']['ms_per_call_incl_warmup']], d['check']['ok'])"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -c 500 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_bench_under_ncu.json 2> gpurun_out/${T}_bench_under_ncu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'slice_rows_fused|slice_cols_tma' -s 2 -c 2 \
    -o gpurun_out/${T}_fused -f python scripts/time_fused.py 1e7 4096 64 1 > gpurun_out/${T}_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'detector_affine_kernel|voxel_finalize_kernel' -s 2 -c 2 \
    -o gpurun_out/${T}_stageb -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_stageb.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'detector_affine_brick_kernel' -s 1 -c 1 \
    -o gpurun_out/${T}_brick -f python scripts/time_detector.py 2048 360 generic affine_tma > gpurun_out/${T}_brick.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fullsize.py -x -q -k "single_slice or config4" > gpurun_out/${T}_memcheck.log 2>&1
tail -3 gpurun_out/${T}_memcheck.log
ls -la gpurun_out | grep ${T} | awk '{print $5, $9}'
