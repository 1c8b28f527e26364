#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r2u2_attn_tests.log 2>&1; echo "attention tests rc=$?"; tail -1 gpurun_out/r2u2_attn_tests.log
python tools/bench_attn.py 2>&1 | head -4
: > gpurun_out/r2u2_knobs.log
for xw in 0 4 5; do
  echo "=== GG_ROLL_XW=$xw (0 = default 8)" >> gpurun_out/r2u2_knobs.log
  GG_ROLL_XW=$xw ALGOS=4 CASES=64:64,64:12 VARIANTS=xf+skip+stats,xf+cat python tools/bench_conv.py >> gpurun_out/r2u2_knobs.log 2>&1
done
cat gpurun_out/r2u2_knobs.log
for pr in 1 0; do
  GG_HALO_PAIR=$pr timeout 600 python bench.py --workload ldm_cfg3 --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2u2_cfg3_pair$pr.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2u2_cfg3_pair$pr.json')); print('cfg3 GG_HALO_PAIR=$pr', round(d['ms_per_step'],3))"
done
