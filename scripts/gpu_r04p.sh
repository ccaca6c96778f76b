#!/bin/bash
mkdir -p gpurun_out
T=r04p
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${T}_pytest.log
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
tail -8 gpurun_out/${T}_trace_e2e.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1.json')); print(round(d['value']), d['e2e']['value'], d['e2e']['seconds_stage_a'], d['e2e']['seconds_stage_b'], d['check']['ok'])"
