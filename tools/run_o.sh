mkdir -p gpurun_out
N=$1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --tb=short -k "$N" > gpurun_out/r2_multi_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_multi_n$N.log; tail -4 gpurun_out/r2_multi_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])
print(json.dumps(d.get('slab')))
PY
