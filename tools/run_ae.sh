#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_tests.log
tail -3 gpurun_out/r2k_tests.log
for w in ldm_cfg3 ldm_cfg4 ccdm_cfg2; do
 for x in 8 0; do
  GG_HALO_XW=$x timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2k_${w}_x$x.json 2> gpurun_out/r2k_${w}_x$x.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2k_${w}_x$x.json")); print("$w GG_HALO_XW=$x", round(d["ms_per_step"],3), round(d["e2e"]["value"],2), round(d["roofline"]["whole_step_frac"],3))
except Exception as e: print("$w $x FAILED", e)
P
 done
done
