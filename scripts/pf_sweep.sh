# L2 prefetch distances (in % of the resident CTAs, 0 = off) of the column pass and of the row passes; side stream on / off
for p in 25 35 50 75; do echo "cols prefetch pct $p"; B4D_COLS_PREFETCH=$p bash scripts/quick_bench.sh "128 0"; done
for p in 25 50 200; do echo "rows prefetch pct $p"; B4D_ROWS_PREFETCH=$p bash scripts/quick_bench.sh "128 0"; done
echo "side stream off"; B4D_SIDE_STREAM=0 bash scripts/quick_bench.sh "128 0"
