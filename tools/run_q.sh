mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2q_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_tests.log; tail -6 gpurun_out/r2q_tests.log
for fx in 0 1; do
  for wl in ldm_cfg3 ldm_cfg4; do
    GG_FUSED_GN_HALO=$fx python bench.py --workload $wl --steps 20 --no-cpu-baseline > gpurun_out/r2q_${wl}_h$fx.json 2>/dev/null; echo "$wl halo-fused=$fx rc=$?"
  done
  GG_FUSED_GN_HALO=$fx python bench.py --no-extras --no-cpu-baseline --steps 8 > gpurun_out/r2q_cfg2_h$fx.json 2>/dev/null; echo "cfg2 halo-fused=$fx rc=$?"
done
python - <<'PY'
import json
for w in ('ldm_cfg3','ldm_cfg4','cfg2'):
    for fx in (0,1):
        d=json.loads(open(f'gpurun_out/r2q_{w}_h{fx}.json').read().strip().splitlines()[-1])
        print(w,'halo-fused',fx,'ms/step %.3f'%d['ms_per_step'],'launches/step',d['gpu_launches']//d['steps'], 'whole', round(d['roofline']['whole_step_frac'],3))
PY
