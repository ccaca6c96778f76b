#!/bin/bash
mkdir -p gpurun_out
T=r04r
for k in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1_$k.json 2> gpurun_out/${T}_bench_n1_$k.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1_$k.json')); print(round(d['value']), round(d['e2e']['value']), d['e2e']['ms_per_call_incl_warmup'], d['check']['ok'])"
done
