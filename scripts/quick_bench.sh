#!/bin/bash
# usage: quick_bench.sh "<frames> <batch>" ...
for cfg in "$@"; do
  set -- $cfg
  python bench.py --steps 5 --warmup 3 --frames $1 --batch $2 --no-cpu-baseline --no-e2e 2>gpurun_out/qb.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('frames',$1,'batch',$2,'fps %.0f'%d['value'],'ms/step %.2f'%d['ms_per_step'],'launches',d['gpu_launches'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
" || tail -5 gpurun_out/qb.err
done
