#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VARIANTS=stats ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 3 -c 1 -o gpurun_out/r2c_upconv_stats python tools/bench_upconv.py > gpurun_out/r2c_ncu_up.log 2>&1; echo "up rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 12 -c 1 -o gpurun_out/r2c_attention_tc python tools/bench_attn.py > gpurun_out/r2c_ncu_attn.log 2>&1; echo "attn rc=$?"
