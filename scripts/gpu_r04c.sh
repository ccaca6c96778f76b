#!/bin/bash
# Round 2, call C (2 GPUs): two-plane fixed-point row kernel, C-ABI NCCL path, T9 on hardware.
mkdir -p gpurun_out
T=r04c
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?"
tail -c 300 gpurun_out/${T}_bench_n1.err
GIWAXS_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 rc=$?"
grep -E "trace|Error|error" gpurun_out/${T}_bench_n2.err | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'slice_rows_fused' -s 1 -c 1 \
    -o gpurun_out/${T}_rows -f python scripts/time_fused.py 1e7 4096 64 1 > gpurun_out/${T}_rows.log 2>&1
ls -la gpurun_out | grep ${T}
