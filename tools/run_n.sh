mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q --tb=short > gpurun_out/r2n_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_tests.log; tail -3 gpurun_out/r2n_tests.log
export ALGOS=4
# ---- launch list of two default bench steps (graph replays included)
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_ccdm_cfg2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_ncu1.log 2>&1; echo "launch list rc=$?"
python bench.py --workload ldm_cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_ldm_cfg3.csv python bench.py --workload ldm_cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_ncu3.log 2>&1; echo "launch list cfg3 rc=$?"
# ---- full captures
CASES=64:12 VARIANTS=xf+cat python tools/bench_conv.py > gpurun_out/r2n_p1.log 2>&1 && \
CASES=64:12 VARIANTS=xf+cat ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 3 -c 1 -o gpurun_out/r2_head_sampler python tools/bench_conv.py > gpurun_out/r2n_ncu_head.log 2>&1; echo "head rc=$?"
CASES=64:64 VARIANTS=xf+res+stats,xf+skip+stats python tools/bench_conv.py > gpurun_out/r2n_p2.log 2>&1 && \
CASES=64:64 VARIANTS=xf+res+stats,xf+skip+stats ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 3 -c 1 -o gpurun_out/r2_roll_res python tools/bench_conv.py > gpurun_out/r2n_ncu_res.log 2>&1; echo "res rc=$?"
CASES=64:64 VARIANTS=xf+skip+stats ncu --set full --clock-control none --import-source on -k regex:conv_roll_kernel -s 3 -c 1 -o gpurun_out/r2_roll_skip python tools/bench_conv.py > gpurun_out/r2n_ncu_skip.log 2>&1; echo "skip rc=$?"
python tools/bench_attn.py > gpurun_out/r2n_p3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 16 -c 1 -o gpurun_out/r2_attention_tc python tools/bench_attn.py > gpurun_out/r2n_ncu_attn.log 2>&1; echo "attn rc=$?"
python tools/prof_pervoxel.py > gpurun_out/r2n_p4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cat_posterior_kernel -s 2 -c 1 -o gpurun_out/r2_cat_posterior python tools/prof_pervoxel.py > gpurun_out/r2n_ncu_cat.log 2>&1; echo "cat rc=$?"
ls -la gpurun_out/*.ncu-rep | tail
