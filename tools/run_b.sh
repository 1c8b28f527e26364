mkdir -p gpurun_out
python -m pytest tests/test_gpu_slab.py tests/test_gpu_models.py -m gpu -q -rA --tb=short -k "slab or tiny_unet or noise_key" > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_tests.log
export CASES=64:64 ALGOS=4
for cfg in "0 0" "1 0" "1 6" "0 5" "0 6" "1 4"; do
  set -- $cfg
  echo "=== SKIP_FIRST=$1 SA=$2" >> gpurun_out/r2b_conv.log
  GG_ROLL_SKIP_FIRST=$1 GG_ROLL_SA=$2 GG_ROLL_DBG=1 VARIANTS=xf+skip+stats python tools/bench_conv.py >> gpurun_out/r2b_conv.log 2>&1
done
echo "=== plain variants" >> gpurun_out/r2b_conv.log
GG_ROLL_DBG=1 VARIANTS=xf+stats,xf+res+stats python tools/bench_conv.py >> gpurun_out/r2b_conv.log 2>&1
echo "=== head" >> gpurun_out/r2b_conv.log
CASES=64:12 GG_ROLL_DBG=1 VARIANTS=xf python tools/bench_conv.py >> gpurun_out/r2b_conv.log 2>&1
