mkdir -p gpurun_out
L=gpurun_out/r2c_conv.log; : > $L
export ALGOS=4
for xw in 4 5 8; do
  echo "=== XW=$xw  64:64 xf+stats,xf+res+stats,xf+skip+stats ; 128:64 xf" >> $L
  GG_ROLL_XW=$xw CASES=64:64 VARIANTS=xf,xf+stats,xf+res+stats,xf+skip+stats python tools/bench_conv.py >> $L 2>&1
done
echo "=== no xform reference: plain,stats,res+stats" >> $L
CASES=64:64 VARIANTS=plain,stats,res+stats,skip+stats python tools/bench_conv.py >> $L 2>&1
for xw in 4 8; do
  echo "=== head XW=$xw xf, xf+cat" >> $L
  GG_ROLL_XW=$xw CASES=64:12 VARIANTS=xf,xf+cat python tools/bench_conv.py >> $L 2>&1
done
echo "=== DBG XW=8 xf+stats / xf+skip+stats" >> $L
GG_ROLL_XW=8 GG_ROLL_DBG=1 CASES=64:64 VARIANTS=xf+stats,xf+skip+stats python tools/bench_conv.py 2>&1 | grep -v "^\[conv_roll\]" >> $L
GG_ROLL_XW=8 GG_ROLL_DBG=1 CASES=64:64 VARIANTS=xf+stats,xf+skip+stats python tools/bench_conv.py 2>&1 | grep "^\[conv_roll\]" | tail -4 >> $L
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "roll or conv" > gpurun_out/r2c_tests.log 2>&1; tail -3 gpurun_out/r2c_tests.log
GG_ROLL_XW=8 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "roll" > gpurun_out/r2c_tests8.log 2>&1; tail -3 gpurun_out/r2c_tests8.log
