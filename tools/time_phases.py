"""Cost of the discriminator phase of the CycleGAN step under CUDA-graph replay: the step timed with 4, 1 and 0-ish
discriminator updates (D_ITERS patched) — (t4 - t1) / 3 is one discriminator update of both PatchGANs."""
import contextlib, io, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel


def run(d_iters, steps=10, **optkw):
    CycleGANModel.D_ITERS = d_iters
    torch.manual_seed(0); random.seed(1234)
    m = CycleGANModel()
    opt = bench.make_opt("cuda", True)
    for k, v in optkw.items():
        setattr(opt, k, v)
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(opt)
    a, b = bench.synthetic_batch(8, 256, 1234)
    dev = {"img_source": a.cuda(), "img_target": b.cuda()}
    for _ in range(6):
        m.set_input(dev); m.optimize_parameters("train")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.set_input(dev); m.optimize_parameters("train")
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


t4 = run(4); t1 = run(1); t2 = run(2)
print("D_ITERS 4: %.2f ms, 2: %.2f ms, 1: %.2f ms -> one discriminator update (both nets) %.2f ms, rest of the step %.2f ms"
      % (t4, t2, t1, (t4 - t1) / 3, t1 - (t4 - t1) / 3))
t4s = run(4, concurrent_D=False)
print("without the two discriminator streams: D_ITERS 4: %.2f ms" % t4s)
