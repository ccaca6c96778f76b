#!/bin/bash
# Round 2, call W (1 GPU): early-exit of the clipped-bbox pass, NVML clock sampler, data-pipe bound in the bench line
mkdir -p gpurun_out
T=r04w
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n1.json')); print(round(d['value']), {k:round(v*1e3,2) for k,v in d['roofline']['kernel_ms_per_slice'].items()}, round(d['e2e']['value']), [c[0] for c in d['e2e']['ms_per_call_incl_warmup']], d['check']['ok'], d['clocks'], d['roofline']['binding_bound'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_bench_under_ncu.json 2> gpurun_out/${T}_bench_under_ncu.err
python scripts/ncu_summary.py gpurun_out/${T}_launches.csv 2>/dev/null | head -30
