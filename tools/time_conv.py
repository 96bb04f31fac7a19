"""Times the dominant convolution shapes (CUDA events, torch current stream)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cycle_depth_estimation_b200 import ops  # noqa: E402


def time_conv(n, c, co, hp, wp, k, stats=True, iters=30, flip=False):
    x = torch.randn((n, hp, wp, c), device="cuda").to(torch.bfloat16)
    w = (torch.randn((co, c, k, k), device="cuda") * 0.02).contiguous()
    wpk, rows_pad, kpad = ops.pack_conv_weight(w, True)
    ho, wo = hp - k + 1, wp - k + 1
    if os.environ.get("PITCHED", "1") == "1":
        y = ops.alloc_flat_output(n, ho, wo, wp, ops.round_up(co, 8), "cuda")
    else:
        y = torch.empty((n, ho, wo, ops.round_up(co, 8)), dtype=torch.bfloat16, device="cuda")
    st = torch.zeros((n, co, 2), dtype=torch.float32, device="cuda") if stats else None
    g = ops.geom(k, k, flip=flip)
    ov = ops.out_view_nhwc(y, co)
    for _ in range(3):
        ops.conv2d_fwd(g, x, wpk, rows_pad, kpad, ov, None, ops.ACT_NONE, 0.0, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.conv2d_fwd(g, x, wpk, rows_pad, kpad, ov, None, ops.ACT_NONE, 0.0, st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    fl = 2.0 * n * ho * wo * c * co * k * k
    print("conv n%d c%d->%d %dx%d k%d stats=%s: %.1f us  %.0f TFLOP/s" % (n, c, co, hp, wp, k, stats, us, fl / us / 1e6),
          flush=True)


if __name__ == "__main__":
    time_conv(8, 256, 256, 66, 66, 3, True)
    time_conv(8, 256, 256, 66, 66, 3, False)
    time_conv(8, 256, 256, 68, 68, 3, False, flip=True)
    time_conv(8, 128, 128, 130, 130, 3, True)
    time_conv(8, 64, 3, 262, 262, 7, False)
    time_conv(8, 256, 512, 34, 34, 4, True)
