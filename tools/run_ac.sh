#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_tests.log
tail -3 gpurun_out/r2h_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2h_smoke.log
for w in ccdm_cfg2 ldm_cfg3 ldm_cfg4 ccdm_cfg1 ccdm_cfg5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-cpu-baseline --detail > gpurun_out/r2h_$w.json 2> gpurun_out/r2h_$w.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2h_$w.json")); print("$w", round(d["ms_per_step"],3), round(d["e2e"]["value"],2), round(d["roofline"]["whole_step_frac"],3), round(d["roofline"]["frac"],3))
except Exception as e: print("$w FAILED", e)
P
  cp gpurun_out/bench_detail_$w.txt gpurun_out/r2h_detail_$w.txt 2>/dev/null
done
