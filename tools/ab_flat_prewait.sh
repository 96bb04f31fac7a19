CDB_FLAT_PAIR=1 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -2
for cfg in "CDB_MMA_PREWAIT=1" "CDB_MMA_PREWAIT=0" "CDB_MMA_PREWAIT=1"; do
  echo "== $cfg"
  env $cfg CDB_FLAT_PAIR=1 python tools/flat_dbg.py 2>&1 | grep "batch 16" | cut -c1-330
done
for cfg in "CDB_MMA_PREWAIT=1" "CDB_MMA_PREWAIT=0"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cudnn-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
done
