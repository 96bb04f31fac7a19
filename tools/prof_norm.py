"""One forward + backward of the norm kernels on the residual-block shape (for an ncu capture)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cycle_depth_estimation_b200 import ops

n, h, w, c, pad = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 64, 64, 256, 1
y = ops.alloc_flat_output(n, h, w, w + 2, c, "cuda"); y.normal_()
stats = torch.zeros((n, c, 2), device="cuda"); ops.channel_stats(y, c, True, stats)
full = torch.empty((n, h + 2 * pad, w + 2 * pad, c), dtype=torch.bfloat16, device="cuda")
inner = full[:, pad:pad + h, pad:pad + w, :]
desc = ops.norm_desc(ops.NORM_INSTANCE, ops.ACT_RELU, 0.0, 1e-5, c, pad, stats)
dfull = torch.randn((n, h + 2 * pad, w + 2 * pad, c), device="cuda").to(torch.bfloat16)
dinner = dfull[:, pad:pad + h, pad:pad + w, :]
dyp = torch.zeros((n, h + 4, w + 4, c), dtype=torch.bfloat16, device="cuda"); dy = dyp[:, 2:2 + h, 2:2 + w, :]
bst = torch.zeros((n, c, 2), device="cuda")
for _ in range(3):
    ops.norm_act_fwd(desc, y, inner, None)
    ops.norm_act_bwd(desc, y, dy, dinner, None, bst, None)
torch.cuda.synchronize()
