# columns per CTA of the column pass (4: two CTAs per SM, 8: one) with and without the PSD map
for cw in 4 8; do for psd in "" "--no-psd"; do
  echo "CW=$cw $psd"
  B4D_COLS_CW=$cw python scripts/sched_sweep.py --steps 10 $psd --configs 0:1:1:0:0:0,0:1:1:0:0:0 2>&1 | tail -1
done; done
