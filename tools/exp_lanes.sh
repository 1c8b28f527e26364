#!/bin/bash
# sample-lane experiment: the lanes test, then each workload at several lane counts (one JSON line each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py -x -q -m gpu -k "lanes" -s > gpurun_out/r2t_lanes_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_lanes_test.log
tail -5 gpurun_out/r2t_lanes_test.log
for l in 1 2 4; do
  GG_LANES=$l python bench.py --workload ldm_cfg3 --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2t_cfg3_l$l.json 2> gpurun_out/r2t_cfg3_l$l.err
  python - <<P
import json
d=json.load(open("gpurun_out/r2t_cfg3_l$l.json")); print("cfg3 lanes $l", d["ms_per_step"], d["e2e"]["value"], d["roofline"]["whole_step_frac"])
P
done
for l in 1 2; do
  GG_LANES=$l python bench.py --workload ldm_cfg4 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t_cfg4_l$l.json 2> gpurun_out/r2t_cfg4_l$l.err
  python - <<P
import json
d=json.load(open("gpurun_out/r2t_cfg4_l$l.json")); print("cfg4 lanes $l", d["ms_per_step"], d["e2e"]["value"], d["roofline"]["whole_step_frac"])
P
done
for l in 1 2; do
  GG_LANES=$l python bench.py --workload ccdm_cfg2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t_cfg2_l$l.json 2> gpurun_out/r2t_cfg2_l$l.err
  python - <<P
import json
d=json.load(open("gpurun_out/r2t_cfg2_l$l.json")); print("cfg2 lanes $l", d["ms_per_step"], d["e2e"]["value"], d["roofline"]["whole_step_frac"])
P
done
