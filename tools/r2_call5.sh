#!/bin/bash
# round 2, GPU call 5 (1 GPU): SELL-U fast path A/B, full GPU suite, the default bench line with its extra configs, C3, C4
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -15
timeout 120 python tools/cycle_probe.py --n 256 --cycles 36 --time-ops | cut -c1-1600
AMGB_SELLU_CTAS=4 timeout 120 python tools/cycle_probe.py --n 256 --cycles 36 --time-ops | cut -c1-1600
( time timeout 600 python bench.py ) > gpurun_out/bench_r2_call5.json 2> gpurun_out/bench_r2_call5.err; tail -c 3500 gpurun_out/bench_r2_call5.json; tail -8 gpurun_out/bench_r2_call5.err
timeout 120 ncu --set full --import-source on --clock-control none -k regex:k_spmv -c 3 -f -o gpurun_out/prof_r2_sellu_fast \
   python tools/cycle_probe.py --n 256 --cycles 1 > gpurun_out/prof_r2_sellu_fast.log 2>&1 || true
( time timeout 400 python bench.py --extras c3 --no-strong --no-async --no-cpu-baseline --steps 2 --warmup 1 ) > gpurun_out/bench_r2_c3.json 2> gpurun_out/bench_r2_c3.err; tail -c 1800 gpurun_out/bench_r2_c3.json; tail -3 gpurun_out/bench_r2_c3.err
( time timeout 540 python bench.py --extras c4 --no-strong --no-async --no-cpu-baseline --steps 1 --warmup 0 ) > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; tail -c 1500 gpurun_out/bench_r2_c4.json; tail -3 gpurun_out/bench_r2_c4.err
ls -la gpurun_out
