mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q --tb=short -x > gpurun_out/r2p_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_tests.log; tail -5 gpurun_out/r2p_tests.log
for fx in 0 1; do
  GG_SPLITK_FIXUP=$fx python bench.py --workload ldm_cfg3 --steps 30 --no-cpu-baseline > gpurun_out/r2p_cfg3_fx$fx.json 2>/dev/null; echo "cfg3 fixup=$fx rc=$?"
  GG_SPLITK_FIXUP=$fx python bench.py --no-extras --no-cpu-baseline --steps 8 > gpurun_out/r2p_cfg2_fx$fx.json 2>/dev/null; echo "cfg2 fixup=$fx rc=$?"
done
python - <<'PY'
import json
for w in ('cfg3','cfg2'):
    for fx in (0,1):
        d=json.loads(open(f'gpurun_out/r2p_{w}_fx{fx}.json').read().strip().splitlines()[-1])
        print(w,'fixup',fx,'ms/step %.3f'%d['ms_per_step'],'launches/step',d['gpu_launches']//d['steps'])
PY
