#!/bin/bash
# final default-bench record of the round (tests / smoke / launch lists: tools/run_ag.sh, tools/run_ao.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 --detail > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
for wl in ccdm_cfg1 ccdm_cfg2_text; do python bench.py --workload $wl --steps 10 --no-cpu-baseline > gpurun_out/r2z_$wl.json 2>/dev/null; echo "$wl rc=$?"; done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2z_smoke.log
