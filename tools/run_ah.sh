#!/bin/bash
# skip-source roll conv: no TMA-store staging tiles (32 KB) -> room for a third 36 KB weight stage next to four plane stages
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2o_skip_sb3.log
for cfg in "0 0" "1 4" "1 5" "1 3"; do
  set -- $cfg
  echo "=== NO_TMA_EPI=$1 SA=$2" >> gpurun_out/r2o_skip_sb3.log
  if [ $1 = 1 ]; then export GG_ROLL_NO_TMA_EPI=1; else unset GG_ROLL_NO_TMA_EPI; fi
  GG_ROLL_SA=$2 ALGOS=4 CASES=64:64 VARIANTS=xf+skip+stats,xf+res+stats python tools/bench_conv.py >> gpurun_out/r2o_skip_sb3.log 2>&1
done
cat gpurun_out/r2o_skip_sb3.log
