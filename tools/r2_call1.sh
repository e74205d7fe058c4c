#!/bin/bash
# round 2, GPU call 1: every GPU test incl. the gated ones, async A/B, iebpx at size, per-operator timings, ncu captures
set -x
mkdir -p gpurun_out
export AMGB_EXPERIMENTAL=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
timeout 700 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -80
timeout 400 python tools/async_fact0_time.py --n 256 --corrections 40 --reps 2
timeout 300 python tools/iebpx_time.py --n 256
timeout 300 python tools/cycle_probe.py --n 256 --cycles 36 --time-ops
timeout 300 python tools/cycle_probe.py --n 256 --cycles 36 --time-ops --explicit
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_async_amg -c 2 -f -o gpurun_out/prof_r2_async \
   python tools/async_fact0_time.py --n 128 --corrections 10 --reps 1 > gpurun_out/prof_r2_async.log 2>&1 || true
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_spmv -c 40 -f -o gpurun_out/prof_r2_cycle \
   python tools/cycle_probe.py --n 256 --cycles 1 > gpurun_out/prof_r2_cycle.log 2>&1 || true
ls -la gpurun_out
