"""CUDA-event timings of the weight-gradient kernel (csrc/conv_wgrad.cu, incl. the split-K finalize) on the layers of a
CycleGAN step: the 3x3 256->256 residual-block layer, the stride-2 / transposed generator layers, the PatchGAN layers."""
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


def case(name, n, ci, co, h, k, stride, pad, transposed=False, outpad=0):
    if transposed:
        x = torch.randn((n, h, h, ci), device='cuda').to(torch.bfloat16)
        ho = (h - 1) * stride - 2 * pad + k + outpad
        dy = torch.randn((n, ho, ho, co), device='cuda').to(torch.bfloat16)
        dw = torch.empty((ci, co, k, k), device='cuda')
        fl = 2.0 * n * h * h * ci * co * k * k
    else:
        x = torch.randn((n, h + 2 * pad, h + 2 * pad, ci), device='cuda').to(torch.bfloat16)  # padding materialised
        ho = (h + 2 * pad - k) // stride + 1
        dy = torch.randn((n, ho, ho, co), device='cuda').to(torch.bfloat16)
        dw = torch.empty((co, ci, k, k), device='cuda')
        fl = 2.0 * n * ho * ho * ci * co * k * k
    g = ops.geom(k, k, stride, 0 if not transposed else pad, 0 if not transposed else pad, 1, transposed)
    us = timeit(lambda: ops.conv2d_wgrad(g, x, dy, dw, False))
    print("%-34s batch %2d: %7.1f us  %6.0f TFLOP/s" % (name, n, us, fl / us / 1e6), flush=True)


for n in (8, 16, 24):
    case("R256 3x3 256->256 @64", n, 256, 256, 64, 3, 1, 1)
    case("d128 3x3 s2 64->128 @256", n, 64, 128, 256, 3, 2, 1)
    case("d256 3x3 s2 128->256 @128", n, 128, 256, 128, 3, 2, 1)
    case("u128 convT 3x3 s2 256->128 @64", n, 256, 128, 64, 3, 2, 1, True, 1)
    case("u64 convT 3x3 s2 128->64 @128", n, 128, 64, 128, 3, 2, 1, True, 1)
for n in (8, 16):
    case("D conv1 4x4 s2 64->128 @128", n, 64, 128, 128, 4, 2, 1)
    case("D conv2 4x4 s2 128->256 @64", n, 128, 256, 64, 4, 2, 1)
    case("D conv3 4x4 s1 256->512 @32", n, 256, 512, 31, 4, 1, 1)
