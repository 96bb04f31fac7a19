"""Times the fused norm kernels on the residual-block shape (8 x 64 x 64 x 256) and the 256x256x64 shape."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cycle_depth_estimation_b200 import ops

def run(n, h, w, c, pad, res, iters=30):
    y = ops.alloc_flat_output(n, h, w, w + 2, c, "cuda"); y.normal_()
    stats = torch.zeros((n, c, 2), device="cuda"); ops.channel_stats(y, c, True, stats)
    full = torch.empty((n, h + 2 * pad, w + 2 * pad, c), dtype=torch.bfloat16, device="cuda")
    inner = full[:, pad:pad + h, pad:pad + w, :]
    r = torch.randn((n, h, w, c), device="cuda").to(torch.bfloat16) if res else None
    desc = ops.norm_desc(ops.NORM_INSTANCE, ops.ACT_NONE if res else ops.ACT_RELU, 0.0, 1e-5, c, pad, stats)
    dfull = torch.randn((n, h + 2 * pad, w + 2 * pad, c), device="cuda").to(torch.bfloat16)
    dinner = dfull[:, pad:pad + h, pad:pad + w, :]
    dyp = torch.zeros((n, h + 4, w + 4, c), dtype=torch.bfloat16, device="cuda"); dy = dyp[:, 2:2 + h, 2:2 + w, :]
    gsum = torch.empty((n, h, w, c), dtype=torch.bfloat16, device="cuda") if res else None
    bst = torch.zeros((n, c, 2), device="cuda")
    def t(fn):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3
    tf = t(lambda: ops.norm_act_fwd(desc, y, inner, r))
    tb = t(lambda: ops.norm_act_bwd(desc, y, dy, dinner, None, bst, gsum))
    mb = n * h * w * c * 2 / 1e6
    print("n%d %dx%dx%d pad%d res=%d: fwd %.1f us (%.0f GB/s), bwd(reduce+apply) %.1f us (%.0f GB/s)" % (
        n, h, w, c, pad, res, tf, mb * (3 if res else 2) / tf * 1e-3 * 1e3, tb, mb * (5 if not res else 6) / tb * 1e-3 * 1e3), flush=True)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
print("batch", B, {k: v for k, v in os.environ.items() if k.startswith("CDB_")}, flush=True)
run(B, 64, 64, 256, 1, False)
run(B, 64, 64, 256, 1, True)
run(B, 64, 64, 256, 0, False)
run(B, 256, 256, 64, 3, False, iters=10)
run(B, 128, 128, 128, 0, False, iters=10)
