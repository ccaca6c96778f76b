#!/bin/bash
mkdir -p gpurun_out
T=r04u
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1.json')); print(round(d['value']), d['ms_per_step_stage_a'], d['ms_per_step_stage_b'], {k:round(v*1e3,2) for k,v in d['roofline']['kernel_ms_per_slice'].items()}, round(d['e2e']['value']), d['e2e']['ms_per_call_incl_warmup'][-2:], d['check']['ok'], d['gpu_launches'])"
timeout 300 python scripts/time_config1.py > gpurun_out/${T}_config1.log 2>&1; tail -2 gpurun_out/${T}_config1.log
