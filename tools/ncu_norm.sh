#!/bin/bash
# ncu captures of the norm backward variants on the residual-block shape (batch $1): full set for the TMA-staged
# kernels (cold cache), and a light warm-cache metric pass for both variants.
B=${1:-16}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
python tools/prof_norm.py $B > gpurun_out/ncu_norm_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:norm_bwd -s 2 -c 4 -f -o gpurun_out/r4c_norm_tma python tools/prof_norm.py $B > gpurun_out/r4c_ncu_full.log 2>&1
for impl in 1 2; do
  CDB_NORM_BWD_IMPL=$impl ncu --metrics $M --clock-control none --cache-control none -k regex:norm_bwd -s 2 -c 4 --csv \
    --log-file gpurun_out/r4c_norm_warm_impl$impl.csv python tools/prof_norm.py $B > /dev/null 2>&1
  CDB_NORM_BWD_IMPL=$impl ncu --metrics $M --clock-control none -k regex:norm_bwd -s 2 -c 4 --csv \
    --log-file gpurun_out/r4c_norm_cold_impl$impl.csv python tools/prof_norm.py $B > /dev/null 2>&1
done
