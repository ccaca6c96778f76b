#!/bin/bash
mkdir -p gpurun_out
T=r04i
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    scripts/trace_e2e_multi.py > gpurun_out/${T}_trace8.log 2>&1; echo rc=$?
grep -E "^rank [07] call" gpurun_out/${T}_trace8.log | head -12
grep -E "cumulative" -A34 gpurun_out/${T}_trace8.log | head -80
