TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
DP_MODEL=segcycle python tools/dp_bn_parity.py 2>&1 | tail -1
DP_MODEL=segcycle $TR tools/dp_bn_parity.py 2>&1 | grep "segcycle\|Error\|error" | tail -3
DP_MODEL=segcycle CDB_BN_SYNC=0 $TR tools/dp_bn_parity.py 2>&1 | grep "segcycle\|Error\|error" | tail -3
python tools/dp_bn_parity.py | tail -1
$TR tools/dp_bn_parity.py 2>&1 | grep "precision\|Error" | tail -2
