for t in 1024 512 513 256; do
  echo "== CDB_METRICS_THREADS=$t"
  CDB_METRICS_THREADS=$t python bench.py --workload metrics --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*' | head -2
done
