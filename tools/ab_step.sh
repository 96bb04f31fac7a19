#!/bin/bash
# A/B of the whole CycleGAN step (bench.py, batch 8, 256x256, graph replay) under environment switches, one box.
# usage: bash tools/ab_step.sh "CDB_X=1" "CDB_X=0 CDB_Y=2" ...   (an empty string "" runs the defaults)
for cfg in "$@"; do
  echo "== ${cfg:-default}"
  env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-cudnn-baseline --no-secondary 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
done
