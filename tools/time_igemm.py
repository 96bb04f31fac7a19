"""CUDA-event timings of the layers that go through the generic implicit-GEMM kernel (csrc/conv_igemm.cu) in a
CycleGAN step: the stride-2 / transposed generator layers and the 4x4 stride-2 PatchGAN layers, forward and data
gradient, with the fused InstanceNorm sums where the step has them."""
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


def case(name, n, ci, co, h, k, stride, pad, transposed, outpad=0, stats=True):
    x = torch.randn((n, h, h, ci), device='cuda').to(torch.bfloat16)
    if transposed:
        w = (torch.randn((ci, co, k, k), device='cuda') * 0.02).contiguous()
        ho = (h - 1) * stride - 2 * pad + k + outpad
    else:
        w = (torch.randn((co, ci, k, k), device='cuda') * 0.02).contiguous()
        ho = (h + 2 * pad - k) // stride + 1
    wp, rows, kpad = ops.pack_conv_weight(w, not transposed)
    y = torch.empty((n, ho, ho, co), dtype=torch.bfloat16, device='cuda')
    st = torch.zeros((n, co, 2), device='cuda') if stats else None
    g = ops.geom(k, k, stride, pad, pad, 1, transposed)
    us = timeit(lambda: ops.conv2d_fwd(g, x, wp, rows, kpad, ops.out_view_nhwc(y, co), None, 0, 0.0, st))
    if transposed:
        fl = 2.0 * n * h * h * ci * co * k * k
    else:
        fl = 2.0 * n * ho * ho * ci * co * k * k
    print("%-34s batch %2d: %7.1f us  %6.0f TFLOP/s" % (name, n, us, fl / us / 1e6), flush=True)


for n in (8, 16, 24):
    case("d128 3x3 s2 64->128 @256", n, 64, 128, 256, 3, 2, 1, False)
    case("d256 3x3 s2 128->256 @128", n, 128, 256, 128, 3, 2, 1, False)
    case("u128 convT 3x3 s2 256->128 @64", n, 256, 128, 64, 3, 2, 1, True, 1)
    case("u64 convT 3x3 s2 128->64 @128", n, 128, 64, 128, 3, 2, 1, True, 1)
    case("dgrad of d128 (convT 128->64 @128)", n, 128, 64, 128, 3, 2, 1, True, 1, stats=False)
    case("dgrad of d256 (convT 256->128 @64)", n, 256, 128, 64, 3, 2, 1, True, 1, stats=False)
    case("dgrad of u128 (s2 128->256 @128)", n, 128, 256, 128, 3, 2, 1, False, stats=False)
    case("dgrad of u64 (s2 64->128 @256)", n, 64, 128, 256, 3, 2, 1, False, stats=False)
for n in (8, 16):
    case("D conv1 4x4 s2 64->128 @128", n, 64, 128, 128, 4, 2, 1, False)
    case("D conv2 4x4 s2 128->256 @64", n, 128, 256, 64, 4, 2, 1, False)
    case("dgrad D conv2 (convT 256->128 @32)", n, 256, 128, 32, 4, 2, 1, True, stats=False)
    case("dgrad D conv1 (convT 128->64 @64)", n, 128, 64, 64, 4, 2, 1, True, stats=False)
