#!/bin/bash
# round 2, GPU call 4 (2 GPUs): every GPU test incl. the 2-GPU ones, then the partitioned bench at N = 2 (weak + strong 512^3),
# graph-captured cycle vs per-operation launches
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 1200 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -40
echo '{"n": 512, "value": 1.0095587768554688, "cycles": 43}' > /tmp/amgb_strong_t1.json     # t(1) measured in call 3 (same hardware, another box)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 1500 $RUN bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 4000 gpurun_out/bench_r2_n2.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_r2_n2.err | tail -15
( time AMGB_DIST_GRAPH=0 timeout 900 $RUN bench.py --gpus 2 --steps 5 --warmup 3 --no-strong ) > gpurun_out/bench_r2_n2_nograph.json 2> gpurun_out/bench_r2_n2_nograph.err; tail -c 2500 gpurun_out/bench_r2_n2_nograph.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_r2_n2_nograph.err | tail -6
ls -la gpurun_out
