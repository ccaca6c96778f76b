#!/bin/bash
# Round 2, call H (8 GPUs): combine variants, direct-DMA shared result, staged slice upload
mkdir -p gpurun_out
T=r04h
run() {  # name, env...
  local name=$1; shift
  env "$@" GIWAXS_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${T}_bench_n8_$name.json 2> gpurun_out/${T}_bench_n8_$name.err; echo "bench $name rc=$?"
  grep -E "trace" gpurun_out/${T}_bench_n8_$name.err | tail -4
}
run allreduce GIWAXS_B200_COMBINE=allreduce
run scatter GIWAXS_B200_COMBINE=scatter
timeout 600 python -m pytest tests/test_gpu_multirank.py -q > gpurun_out/${T}_pytest_multirank.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest_multirank.log
