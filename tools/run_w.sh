#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/bench_upconv.py > gpurun_out/r2a_upconv.log 2>&1; cat gpurun_out/r2a_upconv.log
ALGOS=4 CASES=64:64 VARIANTS=plain,xf+res+stats,xf+skip+stats python tools/bench_conv.py > gpurun_out/r2a_conv.log 2>&1; cat gpurun_out/r2a_conv.log
ALGOS=3 CASES=128:128 VARIANTS=plain,res python tools/bench_conv.py 8 32 64 64 >> gpurun_out/r2a_conv.log 2>&1; tail -2 gpurun_out/r2a_conv.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -3 gpurun_out/r2a_tests.log
for w in ccdm_cfg2 ldm_cfg3 ldm_cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline --detail > gpurun_out/r2a_$w.json 2> gpurun_out/r2a_$w.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2a_$w.json")); print("$w", round(d["ms_per_step"],3), round(d["e2e"]["value"],2), round(d["roofline"]["whole_step_frac"],3), round(d["roofline"]["frac"],3))
except Exception as e: print("$w FAILED", e)
P
  cp gpurun_out/bench_detail_$w.txt gpurun_out/r2a_detail_$w.txt 2>/dev/null
done
