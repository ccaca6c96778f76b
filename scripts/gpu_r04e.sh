#!/bin/bash
# Round 2, call D (2 GPUs): DC sample in fp64, per-row |f| bound kernel, 16384-point schedule, e2e back to normal?
mkdir -p gpurun_out
T=r04e
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?"
tail -c 300 gpurun_out/${T}_bench_n1.err
GIWAXS_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 rc=$?"
grep -E "trace|Error|error" gpurun_out/${T}_bench_n2.err | tail -6
timeout 300 python scripts/time_stage_b.py 360 > gpurun_out/${T}_stage_b_360.log 2>&1
timeout 300 python scripts/time_stage_b.py 45 > gpurun_out/${T}_stage_b_45.log 2>&1
cat gpurun_out/${T}_stage_b_360.log gpurun_out/${T}_stage_b_45.log
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
tail -12 gpurun_out/${T}_trace_e2e.log
ls -la gpurun_out | grep ${T}
