mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_fullsize.py -m gpu -q -rA --tb=short -k "peer or properties" > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_tests.log; tail -6 gpurun_out/r2k_tests.log
