#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "roll or conv" > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2e_tests.log
: > gpurun_out/r2e_sweep.log
for cfg in "1 0" "1 4" "1 3" "1 5" "0 0" "0 4"; do
  set -- $cfg
  echo "=== CW=$1 SA=$2" >> gpurun_out/r2e_sweep.log
  GG_ROLL_CW=$1 GG_ROLL_SA=$2 ALGOS=4 CASES=64:64 VARIANTS=xf+skip+stats,skip+stats python tools/bench_conv.py >> gpurun_out/r2e_sweep.log 2>&1
done
echo "=== CW=1 G=1" >> gpurun_out/r2e_sweep.log
GG_ROLL_CW=1 GG_ROLL_G=1 ALGOS=4 CASES=64:64 VARIANTS=xf+skip+stats python tools/bench_conv.py >> gpurun_out/r2e_sweep.log 2>&1
cat gpurun_out/r2e_sweep.log
