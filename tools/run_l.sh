mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --detail > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err; echo "bench n2 rc=$?"
