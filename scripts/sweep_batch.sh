# frames per internal FFT batch of the fused pipeline: does keeping the row<->column intermediates inside L2 pay?
for b in 2 4 6 8 16 32 128; do
  python bench.py --steps 10 --warmup 3 --frames 128 --batch $b --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('batch',$b,'fps %.0f'%d['value'],'ms/step %.2f'%d['ms_per_step'],'launches',d['gpu_launches'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
"
done
