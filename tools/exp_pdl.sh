#!/bin/bash
# programmatic dependent launch on/off: GPU tests with it on, then each workload both ways
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_tests.log
tail -4 gpurun_out/r2u_tests.log
for w in ldm_cfg3 ldm_cfg4 ccdm_cfg2 ccdm_cfg1; do
 for p in 0 1; do
  GG_PDL=$p timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2u_${w}_pdl$p.json 2> gpurun_out/r2u_${w}_pdl$p.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2u_${w}_pdl$p.json")); print("$w pdl $p", round(d["ms_per_step"],3), round(d["e2e"]["value"],2), round(d["roofline"]["whole_step_frac"],3))
except Exception as e: print("$w pdl $p FAILED", e)
P
 done
done
