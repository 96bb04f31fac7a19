"""One eager CycleGAN training step (BASELINE configs[1]: batch 8, 256x256) after two warm-up steps — the command the
ncu launch list under profiles/ is taken from (every kernel of the step with its device time)."""
import contextlib, io, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cycle_depth_estimation_b200 import _lib
from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel

torch.manual_seed(0); random.seed(1234)
model = CycleGANModel()
with contextlib.redirect_stdout(io.StringIO()):
    model.initialize(bench.make_opt("cuda", False))
a, b = bench.synthetic_batch(8, 256, 1234)
dev = {"img_source": a.cuda(), "img_target": b.cuda()}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n0 = 0
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        n0 = _lib.lib().cdb_launch_count()
    model.set_input(dev); model.optimize_parameters("train")
torch.cuda.synchronize()
print("library launches in the last step:", _lib.lib().cdb_launch_count() - n0, "losses", model.get_current_losses())
