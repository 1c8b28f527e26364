#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ALGOS=4 CASES=64:64 VARIANTS=xf+skip+stats ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 3 -c 1 -o gpurun_out/r2b_roll_skip python tools/bench_conv.py > gpurun_out/r2b_ncu_skip.log 2>&1; echo "skip rc=$?"
ALGOS=4 CASES=64:64 VARIANTS=xf+res+stats ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 3 -c 1 -o gpurun_out/r2b_roll_res python tools/bench_conv.py > gpurun_out/r2b_ncu_res.log 2>&1; echo "res rc=$?"
ALGOS=4 CASES=64:12 VARIANTS=xf+cat ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 3 -c 1 -o gpurun_out/r2b_head python tools/bench_conv.py > gpurun_out/r2b_ncu_head.log 2>&1; echo "head rc=$?"
VARIANTS=stats ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 3 -c 1 -o gpurun_out/r2b_upconv_stats python tools/bench_upconv.py > gpurun_out/r2b_ncu_up.log 2>&1; echo "up rc=$?"
