#!/bin/bash
# final records of the round: smoke, GPU tests, reference arm, default bench (with --detail), secondary workloads, ncu launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2n_smoke.log
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2n_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_tests.log; tail -3 gpurun_out/r2n_tests.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 --detail > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
for wl in ccdm_cfg1 ccdm_cfg2_text; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r2n_$wl.json 2>/dev/null; echo "$wl rc=$?"; done
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2n_launches_ccdm_cfg2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_ncu1.log 2>&1; echo "launch list rc=$?"
python bench.py --workload ldm_cfg3 --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2n_launches_ldm_cfg3.csv python bench.py --workload ldm_cfg3 --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_ncu3.log 2>&1; echo "launch list cfg3 rc=$?"
