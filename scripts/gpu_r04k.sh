#!/bin/bash
# Round 2, call K (8 GPUs): scaling of the final build at 1, 2, 4, 8 ranks (bench.py as the driver runs it)
mkdir -p gpurun_out
T=r04k
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench n1 rc=$?"
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench n$N rc=$?"
done
GIWAXS_B200_TRACE=1 timeout 300 python scripts/trace_config5.py > gpurun_out/${T}_trace_e2e.log 2>&1
tail -6 gpurun_out/${T}_trace_e2e.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    scripts/trace_e2e_multi.py > gpurun_out/${T}_trace8.log 2>&1
grep -E "^rank 0 call" gpurun_out/${T}_trace8.log | head -5
