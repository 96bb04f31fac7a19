for cfg in "X=1" "CDB_IGEMM_KGROUP=1" "CDB_WGRAD_PAIR=0"; do
  echo "== model5 $cfg"; env $cfg python bench.py --workload model5 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
done
echo "== segcycle"; python bench.py --workload segcycle --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
echo "== pix2pix"; python bench.py --workload pix2pix --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
