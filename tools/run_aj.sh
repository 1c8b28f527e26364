#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear or timestep or embedding" > gpurun_out/r2q_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2q_tests.log
for w in ldm_cfg3 ccdm_cfg1 ldm_cfg4; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-extras --no-cpu-baseline --detail > gpurun_out/r2q_$w.json 2> gpurun_out/r2q_$w.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2q_$w.json")); print("$w", round(d["ms_per_step"],3), round(d["e2e"]["value"],1), d["kernel_ms"].get("gg_small_linear"))
except Exception as e: print("$w FAILED", e)
P
done
grep small_linear gpurun_out/bench_detail_ldm_cfg3.txt
