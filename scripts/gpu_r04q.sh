#!/bin/bash
mkdir -p gpurun_out
T=r04q
GIWAXS_B200_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1.json')); print(round(d['value']), d['e2e']['value'], d['e2e']['ms_per_call_incl_warmup'])"
grep "trace. voxel" gpurun_out/${T}_bench_n1.err | cut -c1-220
