#!/bin/bash
# GroupNorm statistics from the halo conv's epilogue: smallest sample (positions) for which it replaces the gg_gn_partial pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in 8192 4096 1024; do
 for w in ldm_cfg3 ldm_cfg4; do
  GG_HALO_STATS_MIN=$m timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2y2_${w}_$m.json 2> gpurun_out/r2y2_${w}_$m.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2y2_${w}_$m.json")); print("$w stats_min $m", round(d["ms_per_step"],3), d["gpu_launches"]//d["steps"])
except Exception as e: print("$w $m FAILED", e)
P
 done
done
