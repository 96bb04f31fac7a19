#!/bin/bash
# A/B of the norm kernel variants (CDB_NORM_BWD_IMPL / CDB_NORM_FWD_IMPL: 1 register loads, 2 TMA-staged) on one box
B=${1:-16}
for v in "CDB_NORM_BWD_IMPL=1 CDB_NORM_FWD_IMPL=1" "CDB_NORM_BWD_IMPL=2" "CDB_NORM_FWD_BLOCKS_PER_SM=3" "CDB_NORM_TMA_HINT=1" "CDB_NORM_TMA_STAGES=2"; do
  echo "== $v"
  env $v timeout 120 python tools/time_norm.py $B 2>&1 | grep -v "^batch"
done
