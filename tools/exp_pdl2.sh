#!/bin/bash
# PDL variants: explicit early trigger (default build) vs completion-only trigger (notrig build), GG_PDL=1; lanes pinned to 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in trig notrig; do
 if [ $v = notrig ]; then export GG_LIB=$PWD/jointimagegeneration_b200/lib/libguidegen_sm100_notrig.so; else unset GG_LIB; fi
 for w in ldm_cfg3 ldm_cfg4 ccdm_cfg1; do
  GG_LANES=1 GG_PDL=1 timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2v_${w}_$v.json 2> gpurun_out/r2v_${w}_$v.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2v_${w}_$v.json")); print("$w $v", round(d["ms_per_step"],3), round(d["e2e"]["value"],2))
except Exception as e: print("$w $v FAILED", e)
P
 done
done
unset GG_LIB
GG_LANES=1 GG_PDL=0 timeout 600 python bench.py --workload ldm_cfg4 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2v_ldm_cfg4_off.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2v_ldm_cfg4_off.json')); print('ldm_cfg4 pdl off lanes 1', round(d['ms_per_step'],3))"
