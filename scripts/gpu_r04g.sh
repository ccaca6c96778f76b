#!/bin/bash
# Round 2, call G (8 GPUs): scaling of the device-timed value and of e2e, with phase traces; T9 at 4 and 8 ranks
mkdir -p gpurun_out
T=r04g
for N in 8 4; do
GIWAXS_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench n$N rc=$?"
grep -E "trace" gpurun_out/${T}_bench_n$N.err | tail -6
done
timeout 600 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/${T}_pytest_multirank.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest_multirank.log
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1; nproc >> gpurun_out/${T}_topo.txt; lscpu | head -20 >> gpurun_out/${T}_topo.txt
