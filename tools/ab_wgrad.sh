for cfg in "CDB_MMA_PREWAIT=1" "CDB_MMA_PREWAIT=0" "CDB_MMA_PREWAIT=1"; do
  echo "== $cfg"
  env $cfg python tools/time_wgrad.py 2>&1 | grep "batch 16\|batch  8" | cut -c1-70
done
