mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rA --tb=short -k "attention" > gpurun_out/r2g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_tests.log; tail -5 gpurun_out/r2g_tests.log
echo "== TC" > gpurun_out/r2g_attn.log; timeout 120 python tools/bench_attn.py >> gpurun_out/r2g_attn.log 2>&1
echo "== mma.sync" >> gpurun_out/r2g_attn.log; GG_ATTN_TC=0 timeout 120 python tools/bench_attn.py >> gpurun_out/r2g_attn.log 2>&1
for kb in 0 384; do
  GG_FUSED_SMALL_GN=1 GG_FUSED_GN_MAX_KB=$kb timeout 200 python bench.py --workload ldm_cfg3 --steps 30 --no-cpu-baseline > gpurun_out/r2g_cfg3_kb$kb.json 2>/dev/null; echo "cfg3 kb=$kb rc=$?"
done
