#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r2t2_attn_tests.log 2>&1; echo "attention tests rc=$?"; tail -3 gpurun_out/r2t2_attn_tests.log
python tools/bench_attn.py > gpurun_out/r2t2_attn.log 2>&1; cat gpurun_out/r2t2_attn.log
