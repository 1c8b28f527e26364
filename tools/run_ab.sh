#!/bin/bash
# knobs measured under the old remote-arrive fence, re-measured: GroupNorm fused into the halo conv, ring shapes of the wide roll conv
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2g_knobs.log
for sa in 0 4; do
  echo "=== roll 64->64 xf+res+stats GG_ROLL_SA=$sa" >> gpurun_out/r2g_knobs.log
  GG_ROLL_SA=$sa ALGOS=4 CASES=64:64,128:64 VARIANTS=xf+res+stats,xf python tools/bench_conv.py >> gpurun_out/r2g_knobs.log 2>&1
done
for xw in 4 8; do
  echo "=== roll xf+res+stats GG_ROLL_XW=$xw" >> gpurun_out/r2g_knobs.log
  GG_ROLL_XW=$xw ALGOS=4 CASES=64:64 VARIANTS=xf+res+stats python tools/bench_conv.py >> gpurun_out/r2g_knobs.log 2>&1
done
cat gpurun_out/r2g_knobs.log
for w in ccdm_cfg2 ldm_cfg3 ldm_cfg4; do
 for f in 0 1; do
  GG_FUSED_GN_HALO=$f timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2g_${w}_h$f.json 2> gpurun_out/r2g_${w}_h$f.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2g_${w}_h$f.json")); print("$w fused_gn_halo $f", round(d["ms_per_step"],3), round(d["roofline"]["whole_step_frac"],3))
except Exception as e: print("$w $f FAILED", e)
P
 done
done
