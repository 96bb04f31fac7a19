"""cProfile of the host side of one CycleGAN step (GPU kept busy; shows where Python time goes)."""
import contextlib, io, os, random, sys, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel
torch.manual_seed(0); random.seed(1234)
model = CycleGANModel()
with contextlib.redirect_stdout(io.StringIO()):
    model.initialize(bench.make_opt("cuda"))
a, b = bench.synthetic_batch(8, 256, 1234)
dev = {"img_source": a.cuda(), "img_target": b.cuda()}
for _ in range(3):
    model.set_input(dev); model.optimize_parameters("train")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    model.set_input(dev); model.optimize_parameters("train")
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue time per step %.1f ms, wall per step %.1f ms" % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
pr = cProfile.Profile()
pr.enable()
model.set_input(dev); model.optimize_parameters("train")
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
