#!/bin/bash
# ncu durations / DRAM bytes of the norm kernels on the residual-block shape (batch $1), register-load (1) vs
# TMA-staged (2) variants, cold cache (ncu default: flush before every replay) and warm (--cache-control none).
B=${1:-16}
TAG=${2:-r4k}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
python tools/prof_norm.py $B > gpurun_out/ncu_norm_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
for impl in 1 2; do
  CDB_NORM_BWD_IMPL=$impl CDB_NORM_FWD_IMPL=$impl ncu --metrics $M --clock-control none --cache-control none -k regex:norm_ -s 3 -c 6 --csv \
    --log-file gpurun_out/${TAG}_norm_warm_impl$impl.csv python tools/prof_norm.py $B > /dev/null 2>&1
  CDB_NORM_BWD_IMPL=$impl CDB_NORM_FWD_IMPL=$impl ncu --metrics $M --clock-control none -k regex:norm_ -s 3 -c 6 --csv \
    --log-file gpurun_out/${TAG}_norm_cold_impl$impl.csv python tools/prof_norm.py $B > /dev/null 2>&1
done
