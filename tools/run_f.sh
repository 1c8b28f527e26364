mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --tb=short > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log; tail -4 gpurun_out/r2f_tests.log
for kb in 0 128 384 1536 49152; do
  GG_FUSED_GN_MAX_KB=$kb python bench.py --workload ldm_cfg3 --steps 30 --no-cpu-baseline > gpurun_out/r2f_cfg3_kb$kb.json 2>/dev/null; echo "cfg3 kb=$kb rc=$?"
done
GG_FUSED_SMALL_GN=0 python bench.py --no-extras --no-cpu-baseline --steps 8 > gpurun_out/r2f_cfg2_nofusedgn.json 2>/dev/null; echo "cfg2 nofused rc=$?"
python bench.py --detail > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
