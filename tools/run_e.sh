mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -rA --tb=short -k "2" > gpurun_out/r2e_multi.log 2>&1; tail -4 gpurun_out/r2e_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r2e_bench_n2.json
