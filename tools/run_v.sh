#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "group_norm" > gpurun_out/r2y_gn_test.log 2>&1; echo "gn pytest rc=$?"; tail -3 gpurun_out/r2y_gn_test.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "group_norm_one_launch" > gpurun_out/r2y_gn_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2y_gn_memcheck.log | tail -3
python tools/bench_gn.py > gpurun_out/r2y_bench_gn.log 2>&1; cat gpurun_out/r2y_bench_gn.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2y_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_tests.log
tail -3 gpurun_out/r2y_tests.log
for w in ldm_cfg3 ccdm_cfg2 ldm_cfg4 ccdm_cfg1; do
 for f in 0 1; do
  GG_FUSED_SMALL_GN=$f timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2y_${w}_f$f.json 2> gpurun_out/r2y_${w}_f$f.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2y_${w}_f$f.json")); print("$w fusedgn $f", round(d["ms_per_step"],3), round(d["e2e"]["value"],2), round(d["roofline"]["whole_step_frac"],3), d["gpu_launches"]//d["steps"])
except Exception as e: print("$w fusedgn $f FAILED", e)
P
 done
done
