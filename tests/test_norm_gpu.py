"""K4 parity at kernel level: statistics, norm + activation (+ residual, + reflect halo) forward and the
fused backward (halo fold + activation mask + norm backward) against torch fp32 autograd on IDENTICAL
bf16 inputs. Tolerance 4e-3 relative L2 (one bf16 rounding of the stored result)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL = 4e-3


def _nhwc(t):  # NCHW float -> NHWC bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _ref_forward(y, norm, act, slope, res, pad, gamma=None, beta=None):
    if norm == 'instance':
        z = F.instance_norm(y, eps=1e-5)
    elif norm == 'batch':
        z = F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5)
    else:
        z = y
    if act == 'relu':
        z = F.relu(z)
    elif act == 'leaky':
        z = F.leaky_relu(z, slope)
    if res is not None:
        z = z + res
    if pad:
        z = F.pad(z, (pad, pad, pad, pad), mode='reflect')
    return z


@pytest.mark.parametrize("norm,act,pad,use_res,c,h,w", [
    ('instance', 'relu', 1, False, 256, 16, 16),
    ('instance', 'none', 1, True, 256, 16, 16),
    ('instance', 'relu', 3, False, 64, 40, 24),
    ('instance', 'leaky', 0, False, 128, 15, 15),
    ('batch', 'relu', 0, False, 72, 8, 12),
    ('batch', 'leaky', 0, False, 1664, 4, 6),
    ('none', 'leaky', 0, False, 64, 20, 20),
    ('none', 'leaky', 0, False, 64, 24, 40),     # rows long enough for the staged backward (PatchGAN first layer)
    ('none', 'relu', 1, False, 128, 12, 64),     # staged, reflect fold, packed ReLU mask from the stored output
])
def test_norm_act_forward_backward(norm, act, pad, use_res, c, h, w):
    from cycle_depth_estimation_b200 import ops
    n = 3
    g = torch.Generator(device='cuda').manual_seed(c + h)
    y = (torch.randn((n, c, h, w), generator=g, device='cuda') * 1.7 + 0.3).to(torch.bfloat16).float()
    res = torch.randn((n, c, h, w), generator=g, device='cuda').to(torch.bfloat16).float() if use_res else None
    gamma = (torch.rand(c, generator=g, device='cuda') + 0.5) if norm == 'batch' else None
    beta = torch.randn(c, generator=g, device='cuda') if norm == 'batch' else None
    yr = y.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if use_res else None
    gr = gamma.clone().requires_grad_(True) if gamma is not None else None
    br = beta.clone().requires_grad_(True) if beta is not None else None
    ref = _ref_forward(yr, norm, act, 0.2, rr, pad, gr, br)
    dout = torch.randn(ref.shape, generator=g, device='cuda').to(torch.bfloat16).float()
    (ref * dout).sum().backward()

    nk = {'instance': ops.NORM_INSTANCE, 'batch': ops.NORM_BATCH, 'none': ops.NORM_NONE}[norm]
    ak = {'relu': ops.ACT_RELU, 'leaky': ops.ACT_LEAKY, 'none': ops.ACT_NONE}[act]
    cs = ops.round_up(c, 8)
    ys = torch.zeros((n, h, w, cs), dtype=torch.bfloat16, device='cuda')
    ys[..., :c] = _nhwc(y)
    stats = None
    if nk != ops.NORM_NONE:
        groups = n if nk == ops.NORM_INSTANCE else 1
        stats = torch.zeros((groups, c, 2), dtype=torch.float32, device='cuda')
        ops.channel_stats(ys, c, nk == ops.NORM_INSTANCE, stats)
        yf = y.double()
        dims = (2, 3) if nk == ops.NORM_INSTANCE else (0, 2, 3)
        s_ref = torch.stack([yf.sum(dims), (yf * yf).sum(dims)], -1).reshape(groups, c, 2)
        assert rel_l2(stats, s_ref) < 1e-5
    full = torch.full((n, h + 2 * pad, w + 2 * pad, cs), float('nan'), dtype=torch.bfloat16, device='cuda')
    inner = full[:, pad:pad + h, pad:pad + w, :]
    rs = None
    if use_res:
        rs = torch.zeros((n, h, w, cs), dtype=torch.bfloat16, device='cuda')
        rs[..., :c] = _nhwc(res)
    desc = ops.norm_desc(nk, ak, 0.2, 1e-5, c, pad, stats, gamma, beta)
    ops.norm_act_fwd(desc, ys, inner, rs)
    got = _nchw(full[..., :c])
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) <= TOL, rel_l2(got, ref)

    # backward: gradient w.r.t. the padded output -> dy (+ gsum for the residual branch)
    dfull = torch.zeros((n, h + 2 * pad, w + 2 * pad, cs), dtype=torch.bfloat16, device='cuda')
    dfull[..., :c] = _nhwc(dout)
    dinner = dfull[:, pad:pad + h, pad:pad + w, :]
    dy = torch.empty((n, h, w, cs), dtype=torch.bfloat16, device='cuda')
    gsum = torch.empty_like(dy) if use_res else None
    groups = n if nk == ops.NORM_INSTANCE else 1
    bstats = torch.zeros((groups, c, 2), dtype=torch.float32, device='cuda')
    yv = ys if nk != ops.NORM_NONE else inner   # norm none: the mask is taken from the stored output
    ops.norm_act_bwd(desc, yv, dy, dinner, None, bstats, gsum)
    assert rel_l2(_nchw(dy[..., :c]), yr.grad) <= TOL, rel_l2(_nchw(dy[..., :c]), yr.grad)
    if use_res:
        assert rel_l2(_nchw(gsum[..., :c]), rr.grad) <= TOL
    if norm == 'batch':
        assert rel_l2(bstats[0, :, 0], br.grad) <= 1e-3
        assert rel_l2(bstats[0, :, 1], gr.grad) <= 1e-3
    if norm == 'none':
        assert rel_l2(bstats[0, :, 0], yr.grad.sum((0, 2, 3))) <= 1e-3


@pytest.mark.parametrize("act,pad,use_res,use_skip,c,h,w,n,ypad_w", [
    ('relu', 1, False, False, 256, 20, 40, 3, 2),     # 16-pixel chunks, 8-pixel tail, y in a wider (flat-conv) frame
    ('none', 1, True, False, 256, 64, 64, 2, 2),      # the residual-block shape (gsum written)
    ('leaky', 0, False, False, 512, 31, 31, 2, 0),    # PatchGAN: 8-pixel chunks, 7-pixel tail
    ('relu', 3, False, True, 64, 37, 200, 2, 0),      # 64-pixel chunks, pad 3 fold, skip gradient added
    ('none', 0, True, True, 128, 24, 72, 5, 0),       # skip + gsum, 32-pixel chunks with an 8-pixel tail
    ('relu', 0, False, False, 2048, 6, 9, 2, 0),      # one pixel per thread group
])
def test_norm_bwd_staged_rows(act, pad, use_res, use_skip, c, h, w, n, ypad_w):
    """The TMA-staged backward (norm_bwd_tma_kernel: InstanceNorm, bf16, contiguous pixel rows): chunk tails, the
    reflect fold at every border, the skip gradient and gsum, against torch fp32 autograd on identical bf16 inputs."""
    from cycle_depth_estimation_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(c + h + w)
    y = (torch.randn((n, c, h, w), generator=g, device='cuda') * 1.3 - 0.2).to(torch.bfloat16).float()
    res = torch.randn((n, c, h, w), generator=g, device='cuda').to(torch.bfloat16).float() if use_res else None
    skip = torch.randn((n, c, h, w), generator=g, device='cuda').to(torch.bfloat16).float() if use_skip else None
    yr = y.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if use_res else None
    z = _ref_forward(yr, 'instance', act, 0.2, rr, 0)
    zp = F.pad(z, (pad, pad, pad, pad), mode='reflect') if pad else z
    dout = torch.randn(zp.shape, generator=g, device='cuda').to(torch.bfloat16).float()
    loss = (zp * dout).sum()
    if use_skip:
        loss = loss + (z * skip).sum()
    loss.backward()

    ak = {'relu': ops.ACT_RELU, 'leaky': ops.ACT_LEAKY, 'none': ops.ACT_NONE}[act]
    yframe = torch.full((n, h, w + ypad_w, c), 3.0, dtype=torch.bfloat16, device='cuda')
    ys = yframe[:, :, :w, :]
    ys.copy_(_nhwc(y))
    stats = torch.zeros((n, c, 2), dtype=torch.float32, device='cuda')
    ops.channel_stats(ys, c, True, stats)
    desc = ops.norm_desc(ops.NORM_INSTANCE, ak, 0.2, 1e-5, c, pad, stats, None, None)
    # forward through the TMA-staged kernel: output + reflect halo, residual read from a wider frame
    full = torch.full((n, h + 2 * pad, w + 2 * pad, c), float('nan'), dtype=torch.bfloat16, device='cuda')
    rs = None
    if use_res:
        rframe = torch.zeros((n, h + 2, w + 4, c), dtype=torch.bfloat16, device='cuda')
        rs = rframe[:, 1:1 + h, 2:2 + w, :]
        rs.copy_(_nhwc(res))
    ops.norm_act_fwd(desc, ys, full[:, pad:pad + h, pad:pad + w, :], rs)
    assert torch.isfinite(full.float()).all()
    assert rel_l2(_nchw(full), zp) <= TOL, rel_l2(_nchw(full), zp)
    dfull = _nhwc(dout)
    dinner = dfull[:, pad:pad + h, pad:pad + w, :]
    dyf = torch.full((n, h + 2, w + 2, c), float('nan'), dtype=torch.bfloat16, device='cuda')
    dy = dyf[:, 1:1 + h, 1:1 + w, :]
    gsum = torch.empty((n, h, w, c), dtype=torch.bfloat16, device='cuda') if use_res else None
    bstats = torch.zeros((n, c, 2), dtype=torch.float32, device='cuda')
    ops.norm_act_bwd(desc, ys, dy, dinner, _nhwc(skip) if use_skip else None, bstats, gsum)
    torch.cuda.synchronize()
    assert torch.isfinite(dy.float()).all()
    assert rel_l2(_nchw(dy), yr.grad) <= TOL, rel_l2(_nchw(dy), yr.grad)
    if use_res:
        assert rel_l2(_nchw(gsum), rr.grad) <= TOL
    assert ops._lib.lib().cdb_device_abort_flag() == 0


def test_nchw_to_nhwc_reflect_and_fold_roundtrip():
    from cycle_depth_estimation_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn((2, 3, 20, 28), generator=g, device='cuda')
    pad = 3
    full = torch.zeros((2, 26, 36, 8), dtype=torch.bfloat16, device='cuda')
    ops.nchw_to_nhwc(x, full[:, pad:23, pad:31, :], pad=pad)
    ref = F.pad(x.to(torch.bfloat16).float(), (pad, pad, pad, pad), mode='reflect')
    assert torch.equal(full[:, :, :34, :3].float().permute(0, 3, 1, 2), ref)
    assert float(full[..., 3:].abs().max()) == 0 and float(full[:, :, 34:, :].abs().max()) == 0
    xr = x.clone().requires_grad_(True)
    up = torch.randn((2, 3, 26, 34), generator=g, device='cuda')
    (F.pad(xr, (pad, pad, pad, pad), mode='reflect') * up).sum().backward()
    out = torch.empty_like(x)
    ops.reflect_fold_nchw(up.contiguous(), out, pad)
    assert rel_l2(out, xr.grad) < 1e-6
