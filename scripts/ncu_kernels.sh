#!/bin/bash
# usage: ncu_kernels.sh <regex> <out-name> [skip] [count]
CMD="python bench.py --steps 1 --warmup 1 --frames 16 --batch 16 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-8} -c ${4:-6} -o gpurun_out/$2 $CMD > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log
