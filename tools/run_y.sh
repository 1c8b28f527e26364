mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log; tail -4 gpurun_out/r2f_tests.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 --detail > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
for wl in ccdm_cfg1 ccdm_cfg2_text; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r2f_$wl.json 2>/dev/null; echo "$wl rc=$?"; done
