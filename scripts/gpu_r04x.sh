#!/bin/bash
# Round 2, call X (8 GPUs): scaling of the FINAL build at 1, 2, 4, 8 ranks (bench.py as the driver runs it),
# the N-rank == 1-rank test through the public drivers, and the per-rank end-to-end trace at 8 ranks
mkdir -p gpurun_out
T=r04x
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?"
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench n$N rc=$?"
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.load(open("gpurun_out/r04x_bench_n%d.json" % n))
        print(n, round(d["value"]), round(d["ms_per_step"], 2), round(d["detector"]["value"]), round(d["e2e"]["value"]),
              round(1e3 * d["e2e"]["seconds_stage_a"], 1), round(1e3 * d["e2e"]["seconds_stage_b"], 2), d["check"]["ok"],
              d["digest"]["count2_crc32"], d["clocks"].get("sm_mhz"))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 900 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/${T}_multirank.log 2>&1; echo "multirank rc=$?"; tail -2 gpurun_out/${T}_multirank.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    scripts/trace_e2e_multi.py > gpurun_out/${T}_trace8.log 2>&1
grep -E "rank 0 call" gpurun_out/${T}_trace8.log | head -5 | cut -c1-200
