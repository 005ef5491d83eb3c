#!/bin/bash
# usage (on the GPU box): scripts/profile_round.sh <tag>
# 1. plain run (must exit 0), 2. ncu launch list of the same command, 3. plain run, 4. ncu --set full of one step's main kernels.
# Outputs under gpurun_out/; scripts/make_profiles.py turns them into profiles/<tag>_*.txt here.
tag=${1:-r01_x}
CMD="python bench.py --steps 2 --warmup 3 --prewarm 0 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { tail -5 gpurun_out/plain_$tag.log; exit 1; }
OURS='cols_kernel|rows_inv|rows_fwd_kernel|rows_moments|frame_reduce|frame_finalize|pilot_kernel|sel_|fused_median|tails_|grain_kernel|argmax_reduce|embed_template|phase_finalize|window_blocks|spec_finalize|f95_|scale_by|absmax'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$OURS" -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'cols_kernel|rows_inv|rows_fwd_kernel|frame_reduce2|fused_median_final|sel_bracket' -s 36 -c 12 \
    -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
