#!/bin/bash
# round 2, GPU call 11 (2 GPUs): the row-partitioned asynchronous solve on the WEAK grid (256^3 rows per GPU), to set its time per
# correction round beside the single-GPU persistent kernel's at 256^3 (16 ms)
set -x
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
( time timeout 200 $RUN bench.py --gpus 2 --steps 3 --warmup 3 --no-strong --async-leg weak ) > gpurun_out/bench_r2_n2_async_weak.json 2> gpurun_out/bench_r2_n2_async_weak.err; tail -c 3000 gpurun_out/bench_r2_n2_async_weak.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_r2_n2_async_weak.err | tail -8
