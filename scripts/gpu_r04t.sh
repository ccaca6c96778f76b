#!/bin/bash
# Round 2, call T (1 GPU): row kernel with the (d, my) table fetched through the texture pipe (GX_F1_TEX=1) vs LDG
mkdir -p gpurun_out
T=r04t
V=giwaxsim_b200/_variants/libgiwaxs_b200_tex.so
GIWAXS_B200_LIB=$V timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest_tex.log 2>&1; echo "pytest(tex) rc=$?"
tail -2 gpurun_out/${T}_pytest_tex.log
for k in 1 2; do
echo "tex:"; GIWAXS_B200_LIB=$V timeout 120 python scripts/time_fused.py 1e7 4096 252 3 2>&1 | tail -1
echo "ldg:"; timeout 120 python scripts/time_fused.py 1e7 4096 252 3 2>&1 | tail -1
done
GIWAXS_B200_LIB=$V timeout 300 python bench.py --no-cpu --no-e2e > gpurun_out/${T}_bench_tex.json 2> gpurun_out/${T}_bench_tex.err; echo "bench(tex) rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_tex.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['roofline']['kernel_ms_per_slice'].items()}, d['check']['ok'])"
GIWAXS_B200_LIB=$V timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:slice_rows_fused -s 2 -c 1 python scripts/time_fused.py 1e7 4096 64 1 2>&1 | grep -E "duration|wavefronts|issue_active" 
