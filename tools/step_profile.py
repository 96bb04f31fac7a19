"""Per-kernel time breakdown of one CycleGAN training step via torch.profiler (CUPTI)."""
import contextlib, io, os, random, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel

torch.manual_seed(0); random.seed(1234)
model = CycleGANModel()
with contextlib.redirect_stdout(io.StringIO()):
    model.initialize(bench.make_opt("cuda"))
a, b = bench.synthetic_batch(8, 256, 1234)
dev = {"img_source": a.cuda(), "img_target": b.cuda()}
for _ in range(2):
    model.set_input(dev); model.optimize_parameters("train")
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    model.set_input(dev); model.optimize_parameters("train")
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name.split("(")[0][:70]
        agg[k][0] += 1; agg[k][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        tot += e.device_time if hasattr(e, "device_time") else e.cuda_time
print("total kernel us", round(tot, 1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%-72s n=%5d %10.1f us %5.1f%% avg %7.1f" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
for kname in ("igemm_flat", "igemm_kernel", "wgrad_kernel", "norm_act_bwd_kernel<true>", "norm_act_fwd"):
    b = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and kname in e.name:
            key = int(round(e.device_time / 5.0) * 5)
            b[key][0] += 1; b[key][1] += e.device_time
    print(kname, sorted([(k, v[0], round(v[1])) for k, v in b.items()], key=lambda t: -t[2])[:14])
