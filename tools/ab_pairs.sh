for cfg in "CDB_WGRAD_PAIR=1 CDB_IGEMM_PAIR=1 CDB_FLAT_PAIR=-1" "CDB_WGRAD_PAIR=0 CDB_IGEMM_PAIR=1" "CDB_WGRAD_PAIR=1 CDB_IGEMM_PAIR=0" "CDB_WGRAD_PAIR=0 CDB_IGEMM_PAIR=0" "CDB_WGRAD_PAIR=0 CDB_IGEMM_PAIR=0 CDB_FLAT_PAIR=0" "CDB_WGRAD_PAIR=1 CDB_IGEMM_PAIR=1 CDB_FLAT_PAIR=-1"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cudnn-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
done
