"""CUDA-event timings of the image-layer kernels (csrc/conv_toeplitz.cu) at the shapes of the CycleGAN step:
c7s1-64 forward / weight gradient and the c7s1-3 data / weight gradient at 256x256, batch 8 / 16 / 24."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cycle_depth_estimation_b200 import ops


def timeit(fn, iters=20):
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


for n in (8, 16, 24):
    h = w = 256
    k = 7
    xb = torch.zeros((n, h + 6, w + 6, 8), dtype=torch.bfloat16, device='cuda')
    xb[..., :3] = torch.randn(n, h + 6, w + 6, 3, device='cuda').to(torch.bfloat16)
    wt = (torch.randn(64, 3, k, k, device='cuda') * 0.05).contiguous()
    wp, rows = ops.pack_toeplitz_weight(wt, True)
    y = ops.alloc_flat_output(n, h, w, w + 6, 64, 'cuda')
    stats = torch.zeros((n, 64, 2), device='cuda')
    t_fwd = timeit(lambda: ops.conv2d_toeplitz_fwd(xb, wp, rows, k, k, ops.out_view_nhwc(y, 64), None, 0, 0.0, stats))
    t_fwd_ns = timeit(lambda: ops.conv2d_toeplitz_fwd(xb, wp, rows, k, k, ops.out_view_nhwc(y, 64)))
    dy = torch.randn(n, h, w, 64, device='cuda').to(torch.bfloat16)
    dw = torch.empty((64, 3, k, k), device='cuda')
    t_wg = timeit(lambda: ops.conv2d_toeplitz_wgrad(dy, xb, k, k, dw, True))
    # output layer: dy 3 channels, zero halo 6 (+8 slack), padded 64-channel input 262 x 262
    dyp = torch.zeros((n, h + 12, w + 12 + 8, 8), dtype=torch.bfloat16, device='cuda')
    dyp[:, 6:6 + h, 6:6 + w, :3] = torch.randn(n, h, w, 3, device='cuda').to(torch.bfloat16)
    w3 = (torch.randn(3, 64, k, k, device='cuda') * 0.05).contiguous()
    wr, rows_r = ops.pack_toeplitz_weight(w3, False, True)
    dfull = ops.alloc_flat_output(n, h + 6, w + 6, dyp.shape[2], 64, 'cuda')
    t_dg = timeit(lambda: ops.conv2d_toeplitz_fwd(dyp, wr, rows_r, k, k, ops.out_view_nhwc(dfull, 64)))
    xpad = torch.randn(n, h + 6, w + 6, 64, device='cuda').to(torch.bfloat16)
    dw3 = torch.empty((3, 64, k, k), device='cuda')
    t_wg3 = timeit(lambda: ops.conv2d_toeplitz_wgrad(xpad, dyp, k, k, dw3, False, flip=True))
    out_mb = n * h * w * 64 * 2 / 1e6
    print("batch %2d: c7s1-64 fwd %.1f us (no stats %.1f; output %.0f MB -> %.0f GB/s), c7s1-64 wgrad %.1f us, "
          "c7s1-3 dgrad %.1f us, c7s1-3 wgrad %.1f us" % (n, t_fwd, t_fwd_ns, out_mb, out_mb / t_fwd * 1e3, t_wg, t_dg, t_wg3))
