for p in 5 10 15 20 25; do echo "cols prefetch pct $p"; B4D_COLS_PREFETCH=$p bash scripts/quick_bench.sh "128 0"; done
