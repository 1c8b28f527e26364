#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2s2_smoke.log
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2s2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s2_tests.log; tail -3 gpurun_out/r2s2_tests.log
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2s2_bench.json')); print('cfg2', round(d['ms_per_step'],2), round(d['e2e']['value'],2), d['clocks']['sm_mhz'])"
