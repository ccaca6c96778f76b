#!/bin/bash
# Round 2, call Z (1 GPU): result download widened by the library's host pool (streaming stores): tests + bench + trace
mkdir -p gpurun_out
T=r04z
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
grep -E "^\[trace\] voxel|^call" gpurun_out/${T}_trace_e2e.log | tail -6
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['roofline']['kernel_ms_per_slice'].items()}, round(d['e2e']['value']), [c for c in d['e2e']['ms_per_call_incl_warmup']], d['check']['ok'])"
