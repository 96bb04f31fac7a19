"""Times the TF32 variant (tcgen05 kind::tf32) of the dominant convolution shapes: forward through the flat and the
generic kernels and the weight gradient (CUDA events, torch current stream)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cycle_depth_estimation_b200 import ops  # noqa: E402


def _time(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def fwd(n, c, co, hp, wp, k, pad=0):
    x = ops.round_tf32_(torch.randn((n, hp, wp, c), device="cuda"))
    w = (torch.randn((co, c, k, k), device="cuda") * 0.02).contiguous()
    wpk, rows_pad, kpad = ops.pack_conv_weight_tf32(w, True)
    ho, wo = hp + 2 * pad - k + 1, wp + 2 * pad - k + 1
    y = torch.empty((n, ho, wo, ops.round_up(co, 4)), dtype=torch.float32, device="cuda")
    g = ops.geom(k, k, 1, pad, pad)
    ov = ops.out_view_nhwc(y, co)
    us = _time(lambda: ops.conv2d_fwd(g, x, wpk, rows_pad, kpad, ov))
    fl = 2.0 * n * ho * wo * c * co * k * k
    print("tf32 fwd n%d c%d->%d %dx%d k%d pad%d: %.1f us  %.0f TFLOP/s" % (n, c, co, hp, wp, k, pad, us, fl / us / 1e6),
          flush=True)


def wgrad(n, c, co, hp, wp, k):
    x = ops.round_tf32_(torch.randn((n, hp, wp, c), device="cuda"))
    ho, wo = hp - k + 1, wp - k + 1
    dy = ops.round_tf32_(torch.randn((n, ho, wo, co), device="cuda"))
    dw = torch.zeros((co, c, k, k), device="cuda")
    g = ops.geom(k, k)
    us = _time(lambda: ops.conv2d_wgrad(g, x, dy, dw, False))
    fl = 2.0 * n * ho * wo * c * co * k * k
    print("tf32 wgrad n%d c%d->%d %dx%d k%d: %.1f us (incl. finalize)  %.0f TFLOP/s" % (n, c, co, hp, wp, k, us, fl / us / 1e6),
          flush=True)


if __name__ == "__main__":
    fwd(8, 256, 256, 66, 66, 3)          # flat kernel
    fwd(8, 256, 256, 64, 64, 3, pad=1)   # generic kernel (TMA zero-fill padding)
    fwd(8, 128, 128, 130, 130, 3)
    fwd(8, 256, 512, 34, 34, 4)
    wgrad(8, 256, 256, 66, 66, 3)
    wgrad(8, 128, 128, 130, 130, 3)
