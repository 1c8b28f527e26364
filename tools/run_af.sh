#!/bin/bash
# attention: fraction of the exponentials on the FMA pipe (degree-3 polynomial) instead of the MUFU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2l_attn_poly.log
for t in "" _poly4 _poly3 _poly2; do
  if [ -z "$t" ]; then unset GG_LIB; else export GG_LIB=$PWD/jointimagegeneration_b200/lib/libguidegen_sm100$t.so; fi
  echo "=== lib$t" >> gpurun_out/r2l_attn_poly.log
  python tools/bench_attn.py >> gpurun_out/r2l_attn_poly.log 2>&1
done
export GG_LIB=$PWD/jointimagegeneration_b200/lib/libguidegen_sm100_poly3.so
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r2l_attn_tests.log 2>&1; echo "poly3 attention tests rc=$?"; tail -2 gpurun_out/r2l_attn_tests.log
cat gpurun_out/r2l_attn_poly.log
