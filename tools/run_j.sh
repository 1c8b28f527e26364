mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --tb=short > gpurun_out/r2j_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_tests.log; tail -4 gpurun_out/r2j_tests.log
python bench.py --detail --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; echo "ref rc=$?"
