#!/bin/bash
# round 2, GPU call 10 (2 GPUs, tight timeouts): the row-partitioned asynchronous solve on two GPUs (tests, then the bench's
# asynchronous leg on the 512^3 problem), and the graph-replay experiment with every NCCL call on ONE stream (no overlap
# branch: the captured cycle is a linear chain)
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_dist.py -m gpu -q -rfEs -k "partitioned or two_gpus" 2>&1 | tail -15 | tee gpurun_out/r2_call10_tests.log
AMGB_DIST_GRAPH=1 AMGB_DIST_OVERLAP=0 timeout 150 python -m pytest tests/test_gpu_dist.py -m gpu -q -rfEs -k "two_gpus_match_global" 2>&1 | tail -12 | tee gpurun_out/r2_call10_graph_linear.log
echo '{"n": 512, "value": 1.0095587768554688, "cycles": 43}' > /tmp/amgb_strong_t1.json     # t(1) measured in call 3 (same hardware, another box)
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
if grep -q "passed" gpurun_out/r2_call10_tests.log && ! grep -q "failed" gpurun_out/r2_call10_tests.log; then
   ( time timeout 480 $RUN bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_r2_n2_async.json 2> gpurun_out/bench_r2_n2_async.err; tail -c 6000 gpurun_out/bench_r2_n2_async.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_r2_n2_async.err | tail -8
fi
if grep -q "2 passed" gpurun_out/r2_call10_graph_linear.log; then
   ( time AMGB_DIST_GRAPH=1 AMGB_DIST_OVERLAP=0 timeout 200 $RUN bench.py --gpus 2 --steps 5 --warmup 3 --no-strong ) > gpurun_out/bench_r2_n2_graph_linear.json 2> gpurun_out/bench_r2_n2_graph_linear.err; tail -c 2500 gpurun_out/bench_r2_n2_graph_linear.json; grep -v "^W\|^\*\*\*" gpurun_out/bench_r2_n2_graph_linear.err | tail -5
fi
ls -la gpurun_out
