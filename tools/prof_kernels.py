"""One launch each of the kernels whose ncu --set full summaries are kept under profiles/ (R256 shapes)."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cycle_depth_estimation_b200 import ops
n, c, hw, k = 8, 256, 64, 3
x = torch.randn((n, hw + 2, hw + 2, c), device="cuda").to(torch.bfloat16)
w = (torch.randn((c, c, k, k), device="cuda") * 0.02).contiguous()
wp, rows_pad, kpad = ops.pack_conv_weight(w, True)
y = ops.alloc_flat_output(n, hw, hw, hw + 2, c, "cuda")
stats = torch.zeros((n, c, 2), device="cuda")
for _ in range(2):
    stats.zero_()
    ops.conv2d_fwd(ops.geom(k, k), x, wp, rows_pad, kpad, ops.out_view_nhwc(y, c), None, ops.ACT_NONE, 0.0, stats)
# wgrad of the same layer
dy = torch.randn((n, hw, hw, c), device="cuda").to(torch.bfloat16)
dw = torch.empty_like(w)
for _ in range(2):
    ops.conv2d_wgrad(ops.geom(k, k), x, dy, dw, False)
# norm forward / backward (InstanceNorm + ReLU, reflect halo 1)
full = torch.empty((n, hw + 2, hw + 2, c), dtype=torch.bfloat16, device="cuda")
inner = full[:, 1:1 + hw, 1:1 + hw, :]
desc = ops.norm_desc(ops.NORM_INSTANCE, ops.ACT_RELU, 0.0, 1e-5, c, 1, stats)
dfull = torch.randn((n, hw + 2, hw + 2, c), device="cuda").to(torch.bfloat16)
dinner = dfull[:, 1:1 + hw, 1:1 + hw, :]
dyp = torch.zeros((n, hw + 4, hw + 4, c), dtype=torch.bfloat16, device="cuda")
bst = torch.zeros((n, c, 2), device="cuda")
for _ in range(2):
    ops.norm_act_fwd(desc, y, inner, None)
    bst.zero_()
    ops.norm_act_bwd(desc, y, dyp[:, 2:2 + hw, 2:2 + hw, :], dinner, None, bst, None)
# depth metrics, 128 images of 375x1242
rng = np.random.default_rng(0)
gt = torch.from_numpy(rng.integers(0, 80, (128, 375, 1242), dtype=np.uint8)).cuda()
pr = torch.from_numpy(rng.integers(0, 256, (128, 375, 1242), dtype=np.uint8)).cuda()
for _ in range(2):
    ops.depth_metrics(gt, pr)
torch.cuda.synchronize()
print("ok")
