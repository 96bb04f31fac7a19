"""The image-layer kernels (csrc/conv_toeplitz.cu): 7x7 convolutions with <= 8 input channels through the Toeplitz
operand, against torch fp32 on the same bf16-rounded operands (relative L2 <= 4e-3, the single-convolution bar of
tests/test_conv_gpu.py), InstanceNorm sums included; forward form (models/networks.py:158) and data-gradient form
of the 3-channel output layer (:185)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def _nhwc8(x_nchw, pad, mode, slack_w=0):
    """fp32 NCHW -> contiguous bf16 [n, h+2p, w+2p+slack, 8] with the padding materialised."""
    xp = F.pad(x_nchw, (pad, pad, pad, pad), mode=mode) if pad else x_nchw
    n, c, h, w = xp.shape
    buf = torch.zeros((n, h, w + slack_w, 8), dtype=torch.bfloat16, device=x_nchw.device)
    buf[:, :, :w, :c] = xp.permute(0, 2, 3, 1).to(torch.bfloat16)
    return buf


@pytest.mark.parametrize("n,h,w,cout,k", [(2, 64, 64, 64, 7), (3, 40, 72, 64, 7), (1, 256, 256, 64, 7), (2, 33, 47, 32, 5),
                                          (2, 32, 32, 128, 3)])
@pytest.mark.parametrize("pitched", [True, False])
def test_toeplitz_forward(n, h, w, cout, k, pitched):
    from cycle_depth_estimation_b200 import ops
    torch.manual_seed(n * 100 + h)
    pad = k // 2
    x = torch.randn(n, 3, h, w, device='cuda')
    wt = (torch.randn(cout, 3, k, k, device='cuda') * 0.05).contiguous()
    bias = torch.randn(cout, device='cuda')
    xb = _nhwc8(x, pad, 'reflect')
    wp, rows_pad = ops.pack_toeplitz_weight(wt, True)
    wpitch = xb.shape[2]
    if pitched:
        y = ops.alloc_flat_output(n, h, w, wpitch, cout, 'cuda')
    else:
        y = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device='cuda')
    stats = torch.zeros((n, cout, 2), dtype=torch.float32, device='cuda')
    ops.conv2d_toeplitz_fwd(xb, wp, rows_pad, k, k, ops.out_view_nhwc(y, cout), bias, ops.ACT_NONE, 0.0, stats)
    ref = F.conv2d(xb[..., :3].permute(0, 3, 1, 2).float(), wt.to(torch.bfloat16).float(), bias)
    got = y.permute(0, 3, 1, 2).float()
    assert got.shape == ref.shape
    assert rel_l2(got, ref) <= 4e-3, rel_l2(got, ref)
    assert rel_l2(stats[..., 0], ref.sum((2, 3)), floor=1.0) <= 2e-3
    assert rel_l2(stats[..., 1], (ref * ref).sum((2, 3))) <= 2e-3
    assert ops._lib.lib().cdb_device_abort_flag() == 0


@pytest.mark.parametrize("n,h,w", [(2, 64, 64), (1, 256, 256), (3, 48, 80)])
def test_toeplitz_data_gradient_of_the_output_layer(n, h, w):
    """c7s1-3 (64 -> 3 channels, reflect pad 3): gradient w.r.t. the PADDED 64-channel input buffer = convolution of the
    zero-haloed (halo 6) 3-channel dy with the flipped filter, written in dy's pitch."""
    from cycle_depth_estimation_b200 import ops
    torch.manual_seed(7)
    k, ci, co = 7, 64, 3
    wt = (torch.randn(co, ci, k, k, device='cuda') * 0.05).contiguous()
    dy = torch.randn(n, co, h, w, device='cuda')
    dyp = _nhwc8(dy, k - 1, 'constant', slack_w=8)          # zero halo of k-1 pixels (+ slack columns, as the engine has)
    hp, wp_ = h + k - 1, w + k - 1                           # the padded input of the forward convolution
    wp, rows_pad = ops.pack_toeplitz_weight(wt, False, flip=True)
    dfull = ops.alloc_flat_output(n, hp, wp_, dyp.shape[2], ci, 'cuda')
    ops.conv2d_toeplitz_fwd(dyp, wp, rows_pad, k, k, ops.out_view_nhwc(dfull, ci))
    xpad = torch.zeros(n, ci, hp, wp_, device='cuda', requires_grad=True)
    out = F.conv2d(xpad, wt.to(torch.bfloat16).float())
    out.backward(dy.to(torch.bfloat16).float())
    got = dfull.permute(0, 3, 1, 2).float()
    assert rel_l2(got, xpad.grad) <= 4e-3, rel_l2(got, xpad.grad)


@pytest.mark.parametrize("n,h,w,cout,k", [(2, 64, 64, 64, 7), (3, 40, 72, 64, 7), (1, 256, 256, 64, 7), (2, 33, 47, 32, 5),
                                          (4, 32, 100, 128, 3)])
def test_toeplitz_weight_gradient_of_the_input_layer(n, h, w, cout, k):
    """c7s1-64 form: dW[o][c][r][s] = sum dy[p, o] * xpad[p + (r, s), c] with dy a plain NHWC tensor."""
    from cycle_depth_estimation_b200 import ops
    torch.manual_seed(3)
    pad = k // 2
    x = torch.randn(n, 3, h, w, device='cuda')
    dy = torch.randn(n, cout, h, w, device='cuda')
    xb = _nhwc8(x, pad, 'reflect')
    dyb = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dw = torch.empty((cout, 3, k, k), dtype=torch.float32, device='cuda')
    ops.conv2d_toeplitz_wgrad(dyb, xb, k, k, dw, True)
    wt = torch.zeros(cout, 3, k, k, device='cuda', requires_grad=True)
    F.conv2d(xb[..., :3].permute(0, 3, 1, 2).float(), wt).backward(dyb.permute(0, 3, 1, 2).float())
    assert rel_l2(dw, wt.grad) <= 2e-3, rel_l2(dw, wt.grad)
    # a strided (pitched) dy view gives the same result
    dyp = ops.alloc_flat_output(n, h, w, xb.shape[2], cout, 'cuda')
    dyp.copy_(dyb)
    dw2 = torch.empty_like(dw)
    ops.conv2d_toeplitz_wgrad(dyp, xb, k, k, dw2, True)
    assert rel_l2(dw2, wt.grad) <= 2e-3
    assert ops._lib.lib().cdb_device_abort_flag() == 0


@pytest.mark.parametrize("n,h,w", [(2, 64, 64), (1, 256, 256), (3, 48, 80)])
def test_toeplitz_weight_gradient_of_the_output_layer(n, h, w):
    """c7s1-3 form: dW[o][i][r][s] = sum_q xpad[q, i] * dyp[q + (6 - r, 6 - s), o] over the padded 64-channel input."""
    from cycle_depth_estimation_b200 import ops
    torch.manual_seed(4)
    k, ci, co = 7, 64, 3
    xpad = torch.randn(n, ci, h + k - 1, w + k - 1, device='cuda')
    dy = torch.randn(n, co, h, w, device='cuda')
    xb = xpad.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dyp = _nhwc8(dy, k - 1, 'constant', slack_w=8)
    dw = torch.empty((co, ci, k, k), dtype=torch.float32, device='cuda')
    ops.conv2d_toeplitz_wgrad(xb, dyp, k, k, dw, False, flip=True)
    wt = torch.zeros(co, ci, k, k, device='cuda', requires_grad=True)
    F.conv2d(xb.permute(0, 3, 1, 2).float(), wt).backward(dy.to(torch.bfloat16).float())
    assert rel_l2(dw, wt.grad) <= 2e-3, rel_l2(dw, wt.grad)
