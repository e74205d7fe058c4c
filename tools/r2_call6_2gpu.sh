#!/bin/bash
# round 2, GPU call 6 (2 GPUs, tight timeouts): partitioned tests on the per-operation path, the graph-replay experiment with the
# warm-up cycle (every replay has a 60 s deadline), bench at N = 2 (weak + strong 512^3)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -rfEs 2>&1 | tail -12
AMGB_DIST_GRAPH=1 timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q -rfEs -k "two_gpus_match_global" 2>&1 | tail -12 | tee gpurun_out/graph_2gpu_test.log
echo '{"n": 512, "value": 1.0095587768554688, "cycles": 43}' > /tmp/amgb_strong_t1.json     # t(1) measured in call 3 (same hardware, another box)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 540 $RUN bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 4500 gpurun_out/bench_r2_n2.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_r2_n2.err | tail -8
if grep -q "2 passed" gpurun_out/graph_2gpu_test.log; then
   ( time AMGB_DIST_GRAPH=1 timeout 200 $RUN bench.py --gpus 2 --steps 5 --warmup 3 --no-strong ) > gpurun_out/bench_r2_n2_graph.json 2> gpurun_out/bench_r2_n2_graph.err; tail -c 2500 gpurun_out/bench_r2_n2_graph.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_r2_n2_graph.err | tail -5
fi
ls -la gpurun_out
