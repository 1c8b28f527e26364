#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel -s 170 -c 6 -o gpurun_out/r2d_cfg3_tc python bench.py --workload ldm_cfg3 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d_ncu_tc.log 2>&1; echo "tc rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 100 -c 4 -o gpurun_out/r2d_cfg3_halo python bench.py --workload ldm_cfg3 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d_ncu_halo.log 2>&1; echo "halo rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 20 -c 2 -o gpurun_out/r2d_cfg3_attn python bench.py --workload ldm_cfg3 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d_ncu_attn.log 2>&1; echo "attn rc=$?"
