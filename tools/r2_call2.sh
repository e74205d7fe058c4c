#!/bin/bash
# round 2, GPU call 2: new persistent kernel + SELL-U: all GPU tests, per-operator timings, async variants, a short bench, ncu
set -x
mkdir -p gpurun_out
free -g | head -2
timeout 900 python -m pytest tests -m gpu -q -rfEs -x 2>&1 | tail -60
timeout 300 python tools/cycle_probe.py --n 256 --cycles 36 --time-ops
timeout 600 python tools/async_time.py --n 256 --corrections 40 --reps 2 --variants default,explicit,no_sellu,global,read_res
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_call2.json 2> gpurun_out/bench_r2_call2.err; tail -c 6000 gpurun_out/bench_r2_call2.json; tail -5 gpurun_out/bench_r2_call2.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_async_amg -c 3 -f -o gpurun_out/prof_r2_async2 \
   python tools/async_time.py --n 128 --corrections 10 --reps 0 --variants default > gpurun_out/prof_r2_async2.log 2>&1 || true
ls -la gpurun_out
