#!/bin/bash
# DRAM bytes of every conv launch of ONE config-2 forward (final code): source of roofline.traffic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2v2_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_roll_kernel|conv_halo_kernel|conv_tcgen05_kernel|splitk" -c 137 --csv --log-file gpurun_out/r2_conv_dram_ccdm_cfg2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2v2_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r2_conv_dram_ccdm_cfg2.csv
