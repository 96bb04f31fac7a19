"""Data-parallel BatchNorm parity: two pix2pix steps on a global batch of 8 (unet_128, 128x128, no
dropout) run (a) by ONE process and (b) sharded over the ranks of a torchrun launch with the per-layer all-reduce of the
BatchNorm sums; (b) must reproduce (a): losses (mean over ranks) and updated weights.

  python tools/dp_bn_parity.py                                   # writes gpurun_out/dp_bn_ref.pt
  torchrun --nproc-per-node 2 ... tools/dp_bn_parity.py          # compares, prints the relative errors
"""
import argparse, contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from cycle_depth_estimation_b200 import ops
from cycle_depth_estimation_b200.pix2pix_model import Pix2PixModel

PREC = os.environ.get("DP_PREC", "tf32x3")
MODEL = os.environ.get("DP_MODEL", "pix2pix")
torch.manual_seed(0)
if MODEL == "segcycle":
    # SegCycle (CycleGAN + four task-network passes with BatchNorm): global batch 4 at 128x128, eager steps, 12 losses
    import random
    import bench
    from cycle_depth_estimation_b200.seg_cycle import SegCycle
    random.seed(1234)
    m = SegCycle()
    o = bench.make_opt("cuda", False)
    o.seg_ngf = 16
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(o)
    g = torch.Generator().manual_seed(5)
    B = 4
    A, Bt = torch.rand((B, 3, 128, 128), generator=g) * 2 - 1, torch.rand((B, 3, 128, 128), generator=g) * 2 - 1
    la = torch.randint(0, 22, (B, 1, 128, 128), generator=g)
    lb = torch.randint(0, 28, (B, 1, 128, 128), generator=g)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    losses = []
    with ops.precision(PREC):
        for _ in range(2):
            m.set_input({'img_source': A[sl].cuda(), 'img_target': Bt[sl].cuda(), 'lab_source': la[sl].cuda(),
                         'lab_target': lb[sl].cuda()})
            m.optimize_parameters('train')
            l = torch.tensor(list(m.get_current_losses().values()), dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(l)
                l /= world
            losses.append(l.cpu())
    path = os.path.join(ROOT, "gpurun_out", "dp_bn_ref_segcycle_%s.pt" % PREC)
    if world == 1:
        torch.save({'losses': losses}, path)
        print("segcycle reference written:", [round(float(x), 5) for x in losses[1]])
    else:
        if rank == 0:
            ref = torch.load(path)
            worst = max(float(((a - b).abs() / b.abs().clamp_min(1e-12)).max()) for a, b in zip(losses, ref['losses']))
            print("segcycle precision %s, world %d, CDB_BN_SYNC=%s: worst relative error of the 12 losses over 2 steps %.3e"
                  % (PREC, world, os.environ.get("CDB_BN_SYNC", "1"), worst))
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
opt = argparse.Namespace(input_nc=3, output_nc=3, ngf=32, ndf=32, netG='unet_128', netD='basic', n_layers_D=3, norm='batch',
                         no_dropout=True, init_type='normal', init_gain=0.02, no_lsgan=True, pool_size=0, lr=2e-4, beta1=0.5,
                         lambda_L1=100.0, isTrain=True, device='cuda', direction='AtoB')
model = Pix2PixModel()
with contextlib.redirect_stdout(io.StringIO()):
    model.initialize(opt)
g = torch.Generator().manual_seed(5)
B = 8
A, Bt = torch.rand((B, 3, 128, 128), generator=g) * 2 - 1, torch.rand((B, 3, 128, 128), generator=g) * 2 - 1
per = B // world
sl = slice(rank * per, (rank + 1) * per)
losses = []
with ops.precision(PREC):
    for _ in range(2):
        model.set_input({'A': A[sl].cuda(), 'B': Bt[sl].cuda(), 'A_paths': None})
        model.optimize_parameters()
        l = torch.tensor(list(model.get_current_losses().values()), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(l)
            l /= world
        losses.append(l.cpu())
state = {k: v.detach().float().cpu() for k, v in list(model.netG.state_dict().items()) + [("D." + k, v) for k, v in model.netD.state_dict().items()]}
path = os.path.join(ROOT, "gpurun_out", "dp_bn_ref_%s.pt" % PREC)
if world == 1:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({'losses': losses, 'state': state}, path)
    print("reference written:", [[round(float(x), 6) for x in l] for l in losses])
elif rank == 0:
    ref = torch.load(path)
    def rel(a, b):
        return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    worst_l = max(float(((a - b).abs() / b.abs().clamp_min(1e-12)).max()) for a, b in zip(losses, ref['losses']))
    errs = {k: rel(state[k], ref['state'][k]) for k in state if state[k].is_floating_point() and 'num_batches' not in k}
    # convolution / transposed-convolution filters (>= 2-D) and BatchNorm scales; the biases in front of a batch
    # normalisation have a mathematically zero gradient, which Adam turns into +-lr noise in BOTH runs (SURVEY B-4)
    filt = {k: v for k, v in errs.items() if state[k].dim() >= 2}
    run_keys = [k for k in errs if 'running_' in k]
    # update of the filters relative to the UPDATE itself (2 Adam steps move every weight by ~2 lr)
    worst_w = max(filt.items(), key=lambda kv: kv[1])
    print("precision %s, world %d, CDB_BN_SYNC=%s: worst relative loss error over 2 steps %.3e; filters: worst relative error "
          "%.3e (%s), median %.3e; running statistics worst %.3e" %
          (PREC, world, os.environ.get("CDB_BN_SYNC", "1"), worst_l, worst_w[1], worst_w[0],
           sorted(filt.values())[len(filt) // 2], max(errs[k] for k in run_keys)))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
