#!/bin/bash
# usage (on the GPU box): scripts/profile_round.sh <tag>
# 1. plain run (must exit 0), 2. ncu launch list of the same command (timed region = NVTX range b4d_timed), 3. plain run,
# 4. ncu --set full of the first step's kernels. Outputs under gpurun_out/; scripts/make_profiles.py turns them into
# profiles/<tag>_*.txt here.
tag=${1:-r02_x}
CMD="python bench.py --steps 2 --warmup 3 --prewarm 0 --frames 128 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "b4d_timed/" -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "b4d_timed/" \
    -k regex:'cols_kernel|rows_inv|rows_fwd_kernel|frame_reduce2|fused_median_final|sel_bracket|temporal_accumulate' --launch-skip 3 -c 10 \
    -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
