mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "halo_fused" 2>&1 | tail -3
for fx in 0 1; do
  for wl in ldm_cfg3 ldm_cfg4; do
    GG_FUSED_GN_HALO=$fx python bench.py --workload $wl --steps 20 --no-cpu-baseline > gpurun_out/r2r_${wl}_h$fx.json 2>/dev/null
  done
  GG_FUSED_GN_HALO=$fx python bench.py --no-extras --no-cpu-baseline --steps 8 --detail > gpurun_out/r2r_cfg2_h$fx.json 2>/dev/null; cp gpurun_out/bench_detail_ccdm_cfg2.txt gpurun_out/r2r_detail_cfg2_h$fx.txt
done
python - <<'PY'
import json
for w in ('ldm_cfg3','ldm_cfg4','cfg2'):
    for fx in (0,1):
        d=json.loads(open(f'gpurun_out/r2r_{w}_h{fx}.json').read().strip().splitlines()[-1])
        print(w,'halo-fused',fx,'ms/step %.3f'%d['ms_per_step'],'launches/step',d['gpu_launches']//d['steps'], 'whole', round(d['roofline']['whole_step_frac'],3), d['kernel_ms'])
PY
