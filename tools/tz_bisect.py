import os, sys
sys.path.insert(0, '/root/repo')
import torch
from cycle_depth_estimation_b200 import ops
n=8; h=w=256; k=7
xb = torch.zeros((n, h + 6, w + 6, 8), dtype=torch.bfloat16, device='cuda')
xb[..., :3] = torch.randn(n, h + 6, w + 6, 3, device='cuda').to(torch.bfloat16)
wt = (torch.randn(64, 3, k, k, device='cuda') * 0.05).contiguous()
wp, rows = ops.pack_toeplitz_weight(wt, True)
y = ops.alloc_flat_output(n, h, w, w + 6, 64, 'cuda')
stats = torch.zeros((n, 64, 2), device='cuda')
use_stats = len(sys.argv) > 1
for _ in range(2):
    ops.conv2d_toeplitz_fwd(xb, wp, rows, k, k, ops.out_view_nhwc(y, 64), None, 0, 0.0, stats if use_stats else None)
torch.cuda.synchronize()
