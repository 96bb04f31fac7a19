#!/bin/bash
# A/B of the norm kernel variants (CDB_NORM_BWD_IMPL / CDB_NORM_FWD_IMPL: 1 register loads, 2 TMA-staged) on one box
# usage: bash tools/ab_norm.sh <batch> ["ENV=.. ENV=.." ...]
B=${1:-16}
shift
if [ $# -eq 0 ]; then set -- "CDB_NORM_BWD_IMPL=1 CDB_NORM_FWD_IMPL=1" "CDB_NORM_BWD_IMPL=2" "CDB_NORM_TMA_V=2"; fi
for v in "$@"; do
  echo "== batch $B ${v:-default}"
  env $v timeout 120 python tools/time_norm.py $B 2>&1 | grep -v "^batch"
done
