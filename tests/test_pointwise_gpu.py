"""K5 / K6b kernels (csrc/pointwise.cu, seg_depth_losses.cu) and the norm-kernel extensions against plain
fp32 torch on identical (bf16-representable) inputs.  Tolerances: the kernels compute in fp32 and round the
result once to bf16 -> relative L2 <= 4e-3; losses (fp32 in / fp32 out) <= 1e-5 relative; masks exact."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu

TOL = 4e-3


def nhwc(n, h, w, c, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((n, h, w, c), generator=g) * scale).to(torch.bfloat16).cuda()


def to_nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def to_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def test_add_and_cast_on_strided_views():
    from cycle_depth_estimation_b200 import ops
    big = nhwc(2, 6, 5, 48, 1)
    a, b = big[..., 8:24], nhwc(2, 6, 5, 16, 2)
    out = torch.empty_like(b)
    ops.add(a, b, out)
    assert rel_l2(out.float(), a.float() + b.float()) <= TOL
    acc = torch.zeros((2, 6, 5, 48), dtype=torch.float32, device="cuda")
    ops.cast(b, acc[..., 16:32], accumulate=True)
    ops.cast(b, acc[..., 16:32], accumulate=True)
    assert torch.equal(acc[..., 16:32], 2 * b.float()) and float(acc[..., :16].abs().max()) == 0.0
    back = torch.empty_like(b)
    ops.cast(acc[..., 16:32], back)
    assert rel_l2(back.float(), 2 * b.float()) <= TOL


def test_avgpool2_forward_backward():
    from cycle_depth_estimation_b200 import ops
    x = nhwc(2, 8, 12, 24, 3)
    out = torch.empty((2, 4, 6, 24), dtype=torch.bfloat16, device="cuda")
    ops.avgpool2_fwd(x, out)
    xr = to_nchw(x).requires_grad_(True)
    ref = F.avg_pool2d(xr, 2, 2)
    assert rel_l2(to_nchw(out), ref) <= TOL
    g = nhwc(2, 4, 6, 24, 4)
    ref.backward(to_nchw(g))
    dx = torch.empty_like(x)
    ops.avgpool2_bwd(g, dx)
    assert rel_l2(to_nchw(dx), xr.grad) <= TOL


@pytest.mark.parametrize("h,w", [(6, 10), (12, 40), (5, 7)])
def test_bilinear2x_align_corners(h, w):
    from cycle_depth_estimation_b200 import ops
    x = nhwc(2, h, w, 16, 5)
    out = torch.empty((2, 2 * h, 2 * w, 16), dtype=torch.bfloat16, device="cuda")
    ops.bilinear2x_fwd(x, out)
    xr = to_nchw(x).requires_grad_(True)
    ref = torch.nn.UpsamplingBilinear2d(scale_factor=2)(xr)
    assert rel_l2(to_nchw(out), ref) <= TOL
    g = nhwc(2, 2 * h, 2 * w, 16, 6)
    ref.backward(to_nchw(g))
    dx = torch.empty_like(x)
    ops.bilinear2x_bwd(g, dx)
    assert rel_l2(to_nchw(dx), xr.grad) <= TOL


def test_attention_gate_forward_backward():
    from cycle_depth_estimation_b200 import ops
    n, h, w, c = 2, 6, 10, 24
    base, s, att = nhwc(n, h, w, c, 7), nhwc(n, h, w, c, 8), nhwc(n, 3, 5, c, 9)
    sums = torch.zeros((n, c, 2), device="cuda")
    ops.channel_stats(att, c, True, sums)
    inv = 1.0 / 15
    out = torch.empty_like(s)
    ops.gate_fwd(base, s, sums, c, inv, out)
    br, sr, ar = (to_nchw(t).requires_grad_(True) for t in (base, s, att))
    ref = br + torch.sigmoid(F.adaptive_avg_pool2d(ar, 1)) * sr
    assert rel_l2(to_nchw(out), ref) <= TOL
    g = nhwc(n, h, w, c, 10)
    ref.backward(to_nchw(g))
    ds = torch.empty_like(s)
    dsum = torch.zeros((n, c), device="cuda")
    ops.gate_bwd(g, s, sums, c, inv, ds, dsum)
    dt = torch.empty_like(att)
    ops.gate_bcast(dsum, sums, c, inv, dt)
    assert rel_l2(to_nchw(ds), sr.grad) <= TOL
    assert rel_l2(to_nchw(dt), ar.grad) <= TOL


def test_prelu_forward_backward():
    from cycle_depth_estimation_b200 import ops
    x = nhwc(2, 9, 7, 16, 11)
    slope = torch.tensor([0.25], device="cuda")
    out = torch.empty_like(x)
    ops.prelu_fwd(x, slope, out)
    xr = to_nchw(x).requires_grad_(True)
    sr = slope.clone().requires_grad_(True)
    ref = F.prelu(xr, sr)
    assert rel_l2(to_nchw(out), ref) <= TOL
    g = nhwc(2, 9, 7, 16, 12)
    ref.backward(to_nchw(g))
    dx = torch.empty_like(x)
    ds = torch.zeros((1,), device="cuda")
    ops.prelu_bwd(x, g, slope, dx, ds)
    assert rel_l2(to_nchw(dx), xr.grad) <= TOL
    assert rel_l2(ds, sr.grad) <= 1e-4


def test_dropout_keeps_half_and_regenerates_mask():
    from cycle_depth_estimation_b200 import ops
    x = torch.ones((4, 16, 16, 64), dtype=torch.bfloat16, device="cuda")
    a, b = torch.empty_like(x), torch.empty_like(x)
    ops.dropout(x, a, 12345, 0.5)
    ops.dropout(x, b, 12345, 0.5)
    assert torch.equal(a, b)
    vals = set(a.float().unique().tolist())
    assert vals == {0.0, 2.0}
    keep = float((a > 0).float().mean())
    assert abs(keep - 0.5) < 0.02
    ops.dropout(x, b, 54321, 0.5)
    assert not torch.equal(a, b)


def test_nhwc_to_nchw_output_conversion():
    from cycle_depth_estimation_b200 import ops
    x = nhwc(2, 5, 9, 32, 13)
    dst = torch.empty((2, 28, 5, 9), device="cuda")
    ops.nhwc_to_nchw(x, 28, dst)
    assert torch.equal(dst, to_nchw(x)[:, :28])


def test_cross_entropy_2d_with_ignore_index():
    from cycle_depth_estimation_b200 import ops
    g = torch.Generator().manual_seed(14)
    logits = (torch.randn((2, 28, 12, 20), generator=g) * 3).cuda()
    labels = torch.randint(0, 28, (2, 12, 20), generator=g)
    labels[torch.rand((2, 12, 20), generator=g) < 0.1] = 255
    labels = labels.cuda()
    acc = torch.zeros((2,), device="cuda")
    grad = torch.empty_like(logits)
    ops.loss_ce2d(logits, labels, 255, acc, grad)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, ignore_index=255)
    ref.backward()
    assert int(acc[1]) == int((labels != 255).sum())
    assert abs(float(acc[0] / acc[1]) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_l2(grad / acc[1], lr.grad) <= 1e-5


def test_bcedep_loss_and_masks():
    from cycle_depth_estimation_b200 import ops
    g = torch.Generator().manual_seed(15)
    x = torch.tanh(torch.randn((2, 1, 10, 16), generator=g)).cuda()
    t = (torch.rand((2, 4, 10, 16), generator=g) * 2 - 1)
    t[t > 0.8] = 1.0
    t[t < -0.8] = -1.0
    t = t.cuda()
    loss = torch.zeros((), device="cuda")
    grad = torch.empty_like(x)
    ops.loss_bcedep(x, t, 50.0, loss, grad)
    xr = x.clone().requires_grad_(True)
    o_m, z_m = (t == 1).float(), (t == -1).float()
    ref = (F.binary_cross_entropy((xr + 1) / 2 * o_m, (t + 1) / 2 * o_m)
           + F.binary_cross_entropy((xr + 1) / 2 * z_m, (t + 1) / 2 * z_m) + 50 * F.l1_loss(xr.expand_as(t), t))
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_l2(grad, xr.grad) <= 1e-5


def test_conv_epilogue_batch_statistics():
    from cycle_depth_estimation_b200 import ops
    x = nhwc(3, 12, 12, 64, 16)
    w = (torch.randn((32, 64, 3, 3), generator=torch.Generator().manual_seed(17)) * 0.05).cuda()
    wp, rows_pad, kpad = ops.pack_conv_weight(w, True)
    buf = torch.zeros((3, 12, 12, 96), dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros((1, 96, 2), device="cuda")
    dst = buf[..., 64:96]
    ops.conv2d_fwd_ex(ops.geom(3, 3, 1, 1, 1), x, wp, rows_pad, kpad, ops.out_view_nhwc(dst, 32), None, ops.ACT_NONE,
                      0.0, stats[:, 64:96], stats_batch=True)
    ref = F.conv2d(to_nchw(x), w.to(torch.bfloat16).float(), padding=1)
    assert rel_l2(to_nchw(dst), ref) <= TOL
    assert float(buf[..., :64].abs().max()) == 0.0
    assert rel_l2(stats[0, 64:, 0], ref.sum((0, 2, 3))) <= 2e-3
    assert rel_l2(stats[0, 64:, 1], (ref * ref).sum((0, 2, 3))) <= 2e-3
    assert float(stats[0, :64].abs().max()) == 0.0


def test_norm_backward_act_first_and_fp32_accumulate():
    """conv -> LeakyReLU -> BatchNorm ordering (ACT_FIRST) and the fp32 accumulating output (ACCUM_F32)."""
    from cycle_depth_estimation_b200 import ops
    n, h, w, c = 2, 6, 8, 16
    pre = nhwc(n, h, w, c, 18)                           # raw convolution output
    a = F.leaky_relu(pre.float(), 0.02).to(torch.bfloat16)  # what the conv epilogue stores
    gamma = (torch.rand(c, generator=torch.Generator().manual_seed(19)) + 0.5).cuda()
    beta = torch.zeros(c, device="cuda")
    stats = torch.zeros((1, c, 2), device="cuda")
    ops.channel_stats(a, c, False, stats)
    out = torch.empty_like(a)
    desc = ops.norm_desc(ops.NORM_BATCH, ops.ACT_LEAKY, 0.02, 1e-5, c, 0, stats, gamma, beta,
                         flags=ops.NORM_FLAG_ACT_FIRST)
    ops.norm_act_fwd(desc, a, out, None)
    pr = to_nchw(pre).requires_grad_(True)
    ar = F.leaky_relu(pr, 0.02)
    ar_q = ar + (to_nchw(a) - ar).detach()
    ref = F.batch_norm(ar_q, None, None, gamma, beta, True, 0.1, 1e-5)
    assert rel_l2(to_nchw(out), ref) <= TOL
    g = nhwc(n, h, w, c, 20)
    ref.backward(to_nchw(g))
    bstats = torch.zeros((1, c, 2), device="cuda")
    dy = torch.empty_like(a)
    ops.norm_act_bwd(desc, a, dy, g, None, bstats, None)
    assert rel_l2(to_nchw(dy), pr.grad) <= TOL
    # accumulate twice into an fp32 slice
    acc = torch.zeros((n, h, w, 32), dtype=torch.float32, device="cuda")
    desc2 = ops.norm_desc(ops.NORM_BATCH, ops.ACT_LEAKY, 0.02, 1e-5, c, 0, stats, gamma, beta,
                          flags=ops.NORM_FLAG_ACT_FIRST | ops.NORM_FLAG_ACCUM_F32)
    for _ in range(2):
        bstats.zero_()
        ops.norm_act_bwd(desc2, a, acc[..., 8:24], g, None, bstats, None)
    assert rel_l2(to_nchw(acc[..., 8:24]), 2 * pr.grad) <= 1e-3
    assert float(acc[..., :8].abs().max()) == 0.0 and float(acc[..., 24:].abs().max()) == 0.0
