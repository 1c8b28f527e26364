#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/bench_upconv.py > gpurun_out/r2w_upconv.log 2>&1; cat gpurun_out/r2w_upconv.log
python tools/bench_upconv.py 8 16 32 32 128 >> gpurun_out/r2w_upconv.log 2>&1; tail -2 gpurun_out/r2w_upconv.log
VARIANTS=stats ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 3 -c 1 -o gpurun_out/r2_upconv_stats python tools/bench_upconv.py > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
VARIANTS=plain ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 3 -c 1 -o gpurun_out/r2_upconv_plain python tools/bench_upconv.py > gpurun_out/r2w_ncu2.log 2>&1; echo "ncu rc=$?"
