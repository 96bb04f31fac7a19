TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for P in tf32x3 bf16; do
  DP_PREC=$P python tools/dp_bn_parity.py
  DP_PREC=$P $TR tools/dp_bn_parity.py 2>&1 | grep "precision"
  DP_PREC=$P CDB_BN_SYNC=0 $TR tools/dp_bn_parity.py 2>&1 | grep "precision"
done
$TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-cudnn-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*' | head -2
$TR bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-cudnn-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*' | head -2
python bench.py --workload pix2pix --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
$TR bench.py --gpus 2 --workload pix2pix --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"parallelism": "[^"]*"' | head -2
