#!/bin/bash
# Launch list + full-set capture of the main kernels of the fused pipeline (1 GPU, short run).
set -x
CMD="python bench.py --steps 1 --warmup 1 --frames 8 --batch 8 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'frame_reduce|cols_kernel|rows_inv|rows_fwd|sel_hist' -s 40 -c 14 -o gpurun_out/prof $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
