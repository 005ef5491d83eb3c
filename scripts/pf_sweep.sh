for p in 0 50 100 150 200 300; do echo "pct $p"; B4D_COLS_PREFETCH=$p bash scripts/quick_bench.sh "128 0"; done
