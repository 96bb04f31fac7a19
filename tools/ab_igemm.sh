for cfg in "X=0" "CDB_IGEMM_KGROUP=1" "CDB_IGEMM_KGROUP=1 CDB_IGEMM_SMEM_KB=168" "CDB_IGEMM_SMEM_KB=168" "X=0"; do
  echo "== $cfg"
  env $cfg python tools/time_igemm.py 2>&1 | grep "batch 16" | cut -c1-70
done
