mkdir -p gpurun_out
L=gpurun_out/r2d_conv.log; : > $L
export ALGOS=4 CASES=64:64
for cfg in "1 1 0" "3 0 0" "3 1 0" "3 0 5" "3 1 4"; do
  set -- $cfg
  echo "=== G=$1 SKIP_FIRST=$2 SA=$3" >> $L
  GG_ROLL_G=$1 GG_ROLL_SKIP_FIRST=$2 GG_ROLL_SA=$3 GG_ROLL_XW=8 VARIANTS=xf+skip+stats,skip+stats python tools/bench_conv.py >> $L 2>&1
  GG_ROLL_G=$1 GG_ROLL_SKIP_FIRST=$2 GG_ROLL_SA=$3 GG_ROLL_XW=8 GG_ROLL_DBG=1 VARIANTS=xf+skip+stats python tools/bench_conv.py 2>&1 | grep "^\[conv_roll\]" | tail -2 >> $L
done
