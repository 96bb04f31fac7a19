"""One or two launches each of the round-2 kernels whose `ncu --set full` summaries are kept under profiles/:
igemm_flat_kernel (R256, batch 8), toeplitz_conv_kernel / toeplitz_wgrad_kernel (c7s1-64, batch 8),
norm_bwd_stream_kernel reduce + apply (R256, batch 16), igemm_kernel (u64 transposed layer, batch 8), wgrad_kernel (R256)."""
import os, sys, torch
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
# the same layer at batch 16: two waves of tiles -> CTA pairs (igemm_flat_kernel<true>)
x16 = torch.randn((16, hw + 2, hw + 2, c), device="cuda").to(torch.bfloat16)
y16 = ops.alloc_flat_output(16, hw, hw, hw + 2, c, "cuda")
stats16 = torch.zeros((16, c, 2), device="cuda")
for _ in range(2):
    stats16.zero_()
    ops.conv2d_fwd(ops.geom(k, k), x16, wp, rows_pad, kpad, ops.out_view_nhwc(y16, c), None, ops.ACT_NONE, 0.0, stats16)
dy = torch.randn((n, hw, hw, c), device="cuda").to(torch.bfloat16)
dw = torch.empty_like(w)
for _ in range(2):
    ops.conv2d_wgrad(ops.geom(k, k), x, dy, dw, False)
# image layer c7s1-64 at 256x256, batch 8
xb = torch.zeros((n, 262, 262, 8), dtype=torch.bfloat16, device="cuda")
xb[..., :3] = torch.randn(n, 262, 262, 3, device="cuda").to(torch.bfloat16)
w7 = (torch.randn(64, 3, 7, 7, device="cuda") * 0.05).contiguous()
tw, trows = ops.pack_toeplitz_weight(w7, True)
y7 = ops.alloc_flat_output(n, 256, 256, 262, 64, "cuda")
st7 = torch.zeros((n, 64, 2), device="cuda")
dy7 = torch.randn(n, 256, 256, 64, device="cuda").to(torch.bfloat16)
dw7 = torch.empty((64, 3, 7, 7), device="cuda")
for _ in range(2):
    ops.conv2d_toeplitz_fwd(xb, tw, trows, 7, 7, ops.out_view_nhwc(y7, 64), None, 0, 0.0, st7)
    ops.conv2d_toeplitz_wgrad(dy7, xb, 7, 7, dw7, True)
# InstanceNorm + ReLU backward of a residual-block layer, batch 16 (reduce + apply)
n2 = 16
y2 = ops.alloc_flat_output(n2, hw, hw, hw + 2, c, "cuda"); y2.normal_()
st2 = torch.zeros((n2, c, 2), device="cuda"); ops.channel_stats(y2, c, True, st2)
desc = ops.norm_desc(ops.NORM_INSTANCE, ops.ACT_RELU, 0.0, 1e-5, c, 1, st2)
dfull = torch.randn((n2, hw + 2, hw + 2, c), device="cuda").to(torch.bfloat16)
dyp = torch.zeros((n2, hw + 4, hw + 4, c), dtype=torch.bfloat16, device="cuda")
bst = torch.zeros((n2, c, 2), device="cuda")
for _ in range(2):
    bst.zero_()
    ops.norm_act_bwd(desc, y2, dyp[:, 2:2 + hw, 2:2 + hw, :], dfull[:, 1:1 + hw, 1:1 + hw, :], None, bst, None)
# u64: ConvTranspose 3x3 s2 128 -> 64 at 128x128, batch 8 (four parity launches of igemm_kernel)
xu = torch.randn((n, 128, 128, 128), device="cuda").to(torch.bfloat16)
wu = (torch.randn((128, 64, 3, 3), device="cuda") * 0.02).contiguous()
wup, ur, uk = ops.pack_conv_weight(wu, False)
yu = torch.empty((n, 256, 256, 64), dtype=torch.bfloat16, device="cuda")
stu = torch.zeros((n, 64, 2), device="cuda")
for _ in range(2):
    ops.conv2d_fwd(ops.geom(3, 3, 2, 1, 1, 1, True), xu, wup, ur, uk, ops.out_view_nhwc(yu, 64), None, 0, 0.0, stu)
torch.cuda.synchronize()
print("ok")
