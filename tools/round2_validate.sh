#!/bin/bash
# One gpurun call that validates and times everything round 1 left unvalidated (written when the GPU budget was spent):
#   gpurun --timeout 900 -- 'bash tools/round2_validate.sh > gpurun_out/round2_validate.log 2>&1'
# 1. the gated tests (device SmoothTransfer, factorised level-0 transfers in the persistent kernel, graph-captured
#    partitioned cycle); 2. A/B timings at 256^3.
set -x
export AMGB_EXPERIMENTAL=1
timeout 300 python -m pytest tests/test_zz_gpu_extended.py tests/test_gpu_dist.py -m gpu -q -k "device_smooth_transfer or async_factorised or graph_captured or nonsymmetric or hybrid_jgs_single_block or async_noinline or afacx_two_sweeps or reference_blocks"
timeout 300 python tools/async_fact0_time.py --n 256 --corrections 40
timeout 200 python tools/iebpx_time.py --n 256
# 3. one ncu capture of the persistent asynchronous kernel (never captured in round 1): is it starved for instructions?
#    (60 616 SASS instructions, profiles/README.md section 9) -- read smsp__warp_issue_stalled_no_instruction*, 
#    l1tex__data_pipe_lsu_wavefronts*, dram__bytes_* from the report here with `ncu -i ... --page raw --csv`
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_async_amg -c 2 -f -o gpurun_out/prof_r2_async \
   python tools/async_fact0_time.py --n 128 --corrections 10 --reps 1 > gpurun_out/prof_r2_async.log 2>&1 || true
