#!/bin/bash
# round 2, GPU call 7 (1 GPU): A/B of the SELL-U kernel variants inside the whole cycle, full GPU suite on the final build, short bench
set -x
mkdir -p gpurun_out
for v in 5 4 6 8; do
   AMGB_SELLU_CTAS=$v timeout 100 python tools/cycle_probe.py --n 256 --cycles 36 --time-ops > gpurun_out/sellu_variant_$v.json 2>/dev/null
   python - <<PY
import json
d = json.load(open("gpurun_out/sellu_variant_$v.json"))
print("variant $v: ms/cycle %.4f  A0 %.4f ms  A0* %.4f ms" % (d["ms_per_cycle"], d["ops"]["A0"]["ms"], d["ops"]["A0*"]["ms"]))
PY
done
timeout 400 python -m pytest tests -m gpu -q -rfEs 2>&1 | tail -12
( time timeout 300 python bench.py --no-strong --no-cpu-baseline --steps 5 --warmup 3 ) > gpurun_out/bench_r2_call7.json 2> gpurun_out/bench_r2_call7.err; tail -c 2500 gpurun_out/bench_r2_call7.json; tail -4 gpurun_out/bench_r2_call7.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2_cycle.csv python tools/cycle_probe.py --n 256 --cycles 2 > /dev/null 2>&1 || true
ls -la gpurun_out
