#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2w2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w2_tests.log
tail -3 gpurun_out/r2w2_tests.log
python tools/bench_upconv.py 2>&1 | tail -2
ALGOS=3 CASES=128:128 VARIANTS=plain,res python tools/bench_conv.py 8 32 64 64 2>&1 | tail -2
for w in ccdm_cfg2 ldm_cfg3 ldm_cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline --detail > gpurun_out/r2w2_$w.json 2> gpurun_out/r2w2_$w.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2w2_$w.json")); print("$w", round(d["ms_per_step"],3), round(d["e2e"]["value"],2), round(d["roofline"]["whole_step_frac"],3), d["clocks"]["sm_mhz"])
except Exception as e: print("$w FAILED", e)
P
done
grep "2x2x2" gpurun_out/bench_detail_ccdm_cfg2.txt | grep "32x64x64" | head -3
