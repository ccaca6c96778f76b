#!/bin/bash
# Round 2, last call (2 GPUs): what the driver runs at round end, on the final commit - GPU suite (with the 2-rank
# equivalence test), smoke, bench at N = 1 and 2, reference arm
mkdir -p gpurun_out
T=r04final
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
for n in ("n1", "n2", "reference"):
    d = json.load(open("gpurun_out/r04final_bench_%s.json" % n))
    print(n, round(d["value"], 1), round(d["ms_per_step"], 2), round(d["e2e"]["value"], 1), d.get("check", {}).get("ok"),
          d.get("clocks"), (d.get("roofline") or {}).get("binding_bound", {}).get("frac"), d["cpu_baseline"]["value"] if "cpu_baseline" in d else None)
PY
