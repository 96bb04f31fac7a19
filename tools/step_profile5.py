"""Per-kernel time breakdown of one seg/depth (model5) or pix2pix training step via torch.profiler (CUPTI)."""
import argparse, contextlib, io, os, random, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
which = sys.argv[1] if len(sys.argv) > 1 else "model5"
torch.manual_seed(0); random.seed(1234)
if which == "model5":
    from cycle_depth_estimation_b200.model5 import Seg_Depth
    model = Seg_Depth()
    with contextlib.redirect_stdout(io.StringIO()):
        model.initialize(argparse.Namespace(lr=2e-4, beta1=0.5, pool_size=50))
    data = {k: v.cuda() for k, v in bench._model5_batch(8, 192, 640, 90).items()}
    step = lambda: (model.set_input(data, 'train'), model.optimize_parameters('train'))
else:
    from cycle_depth_estimation_b200.pix2pix_model import Pix2PixModel
    opt = argparse.Namespace(input_nc=3, output_nc=3, ngf=64, ndf=64, netG='unet_256', netD='basic', n_layers_D=3,
                             norm='batch', no_dropout=False, init_type='normal', init_gain=0.02, no_lsgan=True,
                             pool_size=0, lr=2e-4, beta1=0.5, lambda_L1=100.0, isTrain=True, device='cuda', direction='AtoB')
    model = Pix2PixModel()
    with contextlib.redirect_stdout(io.StringIO()):
        model.initialize(opt)
    a, b = bench.synthetic_batch(16, 256, 1234)
    dev = {'A': a.cuda(), 'B': b.cuda(), 'A_paths': None}
    step = lambda: (model.set_input(dev), model.optimize_parameters())
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name.split("(")[0][:70]
        agg[k][0] += 1; agg[k][1] += e.device_time; tot += e.device_time
print(which, "total kernel us", round(tot, 1), "launches", sum(v[0] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-72s n=%5d %10.1f us %5.1f%% avg %7.1f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
