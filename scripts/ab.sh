#!/bin/bash
# A/B of library builds on one box: scripts/ab.sh <rounds> <name-or-path> ...  ("main" = barc4dip_b200/libb4d.so, other
# names = barc4dip_b200/variants/libb4d_<name>.so from `python -m barc4dip_b200.build --variant <name> -D...`).
# The builds take turns (round-robin) so that clock / thermal drift hits all of them alike.
rounds=$1; shift
for r in $(seq $rounds); do
  for v in "$@"; do
    n=${v%%@*}; e=""; [ "$n" != "$v" ] && e=${v#*@}        # name@ENV=VAL,ENV2=VAL2: environment knobs of that arm
    lib=barc4dip_b200/variants/libb4d_$n.so
    [ "$n" = main ] && lib=barc4dip_b200/libb4d.so
    env ${e//,/ } B4D_LIB=$PWD/$lib python bench.py --steps ${AB_STEPS:-10} --warmup 3 --no-cpu-baseline --no-e2e ${AB_ARGS} 2>gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v'.ljust(22),'ms/step %.3f'%d['ms_per_step'],'fps %.0f'%d['value'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items() if v['ms_per_step']>0.05})
" || tail -3 gpurun_out/ab.err
  done
done
