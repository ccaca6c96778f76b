#!/bin/bash
# Round 2, call R (1 GPU): column kernel with an 11-slot ring (8 before) at N = 4096, 16 slots below: parity + A/B
mkdir -p gpurun_out
T=r04r
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log
for k in 1 2; do
echo "ring 11:"; timeout 300 python scripts/time_fused.py 1e7 4096 252 3 2>&1 | tail -1
echo "ring  8:"; GIWAXS_B200_LIB=giwaxsim_b200/_variants/libgiwaxs_b200_ring8.so timeout 300 python scripts/time_fused.py 1e7 4096 252 3 2>&1 | tail -1
done
echo "N=2048 ring 16:"; timeout 300 python scripts/time_fused.py 2.1e6 2048 128 3 2>&1 | tail -1
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['roofline']['kernel_ms_per_slice'].items()}, round(d['e2e']['value']), d['check']['ok'])"
