#!/bin/bash
# halo-brick conv with fused GroupNorm: eight (default) vs four transform warps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in xw8 xw4; do
 if [ $v = xw4 ]; then export GG_LIB=$PWD/jointimagegeneration_b200/lib/libguidegen_sm100_xw4.so; else unset GG_LIB; fi
 for w in ccdm_cfg2 ldm_cfg3 ldm_cfg4; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j_${w}_$v.json 2> gpurun_out/r2j_${w}_$v.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2j_${w}_$v.json")); print("$w $v", round(d["ms_per_step"],3), round(d["e2e"]["value"],2))
except Exception as e: print("$w $v FAILED", e)
P
 done
done
