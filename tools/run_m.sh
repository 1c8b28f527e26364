mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py -m gpu -q -rA --tb=short > gpurun_out/r2m_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_tests.log; tail -8 gpurun_out/r2m_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --detail > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench_n2.json').read().strip().splitlines()[-1])
print(json.dumps(d['slab']))
PY
