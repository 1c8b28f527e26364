#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear or timestep or embedding" > gpurun_out/r2p_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2p_tests.log
timeout 600 python -m pytest tests/test_gpu_models.py -x -q -m gpu > gpurun_out/r2p_tests2.log 2>&1; echo "models pytest rc=$?"; tail -2 gpurun_out/r2p_tests2.log
for cfg in "16 6" "16 12" "16 4" "8 6" "4 8"; do
  set -- $cfg
  GG_SPLITK_MAX=$1 GG_SPLITK_MIN_KB=$2 timeout 600 python bench.py --workload ldm_cfg3 --steps 30 --warmup 5 --no-extras --no-cpu-baseline --detail > gpurun_out/r2p_cfg3_$1_$2.json 2> gpurun_out/r2p_cfg3_$1_$2.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2p_cfg3_$1_$2.json")); print("cfg3 splitk max $1 minkb $2", round(d["ms_per_step"],3), d["kernel_ms"].get("gg_small_linear"))
except Exception as e: print("cfg3 $1 $2 FAILED", e)
P
done
