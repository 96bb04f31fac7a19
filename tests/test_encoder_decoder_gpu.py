"""the U-Net task network of SegCycle (models/encoder_decoder.py: _UNetEncoder, _UNetDecoder) on
the graph engine vs the fp32 oracle (oracle/encoder_decoder_oracle.py, pinned to the reference's own classes) on
identical name-keyed synthetic weights.

Tolerances: forward <= 2e-2 relative L2 (bf16 path) on the first two encoder levels and max(2e-2, 1.25 x the error of
torch's own bf16 autocast of the oracle) behind the small-batch BatchNorm layers of the centre.  Gradients: with the shared PReLU slope set to 1 ('linear':
the only branching activation becomes the identity in BOTH implementations; BatchNorm, tanh, pooling, the scaled
skips, nearest upsampling and all concatenations stay) every gradient must match fp32 torch to 4e-2 — that pins the
hand-written backward wiring; with the real slope the gradients stay inside the activation-flip envelope of a bf16
forward (see test_networks_gpu.py) with a far smaller median."""
import statistics

import pytest
import torch

from helpers import TOL_BF16, rel_l2, seeded_image, true_fp32
from oracle import encoder_decoder_oracle as OE
from oracle import networks5_oracle as O5

pytestmark = pytest.mark.gpu

LINEAR_TOL = 4e-2
GRAD_FLIP_TOL = 0.35
MEDIAN_TOL = 0.2
SLOPE_ABS_TOL = 1e-2     # |slope gradient error| <= 1e-2 x sum of the magnitudes of its terms (cancellation ~1e4)


def _load(net, seed, slope_key, mode):
    sd = OE.tie_prelu(O5.synth_state_dict(net.state_dict(), seed))
    if mode == "linear":
        sd[slope_key] = torch.ones_like(sd[slope_key])
        OE.tie_prelu(sd)
    net.load_state_dict(sd, strict=True)
    return net.cuda().train(), {k: v.cuda() for k, v in sd.items()}


def _leafs(sd):
    return OE.tie_prelu(O5.leaf_params(sd))


class _AbsSlopePrelu(torch.autograd.Function):
    """F.prelu whose slope 'gradient' is sum |g * min(x, 0)|: the magnitude of the terms the true slope gradient
    sums with cancelling signs (used as the scale of an absolute comparison for that one scalar)."""

    @staticmethod
    def forward(ctx, x, a):
        ctx.save_for_backward(x, a)
        return torch.where(x >= 0, x, a * x)

    @staticmethod
    def backward(ctx, g):
        x, a = ctx.saved_tensors
        return torch.where(x >= 0, g, a * g), (g * x.clamp(max=0)).abs().sum().reshape(1)


def _slope_scales(sd_e, sd_d, x0, loss_fn, monkeypatch):
    import torch.nn.functional as F
    le, ld = _leafs(sd_e), _leafs(sd_d)
    with monkeypatch.context() as m:
        m.setattr(F, "prelu", _AbsSlopePrelu.apply)
        with true_fp32():
            loss_fn(OE.unet_decoder(ld, OE.unet_encoder(le, x0))).backward()
    return float(le[OE.ENC_SLOPE].grad), float(ld[OE.DEC_SLOPE].grad)


def _compare_grads(net, ref_sd, mode, slope_key, slope_scale, env_sd=None):
    named = dict(net.named_parameters())
    gmax = max(float(r.grad.norm()) for r in ref_sd.values() if r.requires_grad and r.grad is not None)
    errs, env = [], []
    for k, p in named.items():
        r = ref_sd[k]
        if r.grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, "missing gradient for " + k
        if k == slope_key:
            # ONE scalar summed over every PReLU of the network with cancelling signs: relative to the net value it
            # is noise-dominated in bf16, so it is compared against the magnitude of its terms
            assert (rel_l2(p.grad, r.grad) <= LINEAR_TOL
                    or abs(float(p.grad) - float(r.grad)) <= SLOPE_ABS_TOL * slope_scale), (k, float(p.grad), float(r.grad), slope_scale)
            continue
        if float(r.grad.norm()) < 1e-3 * gmax:
            # (near-)zero by construction (a bias in front of a batch-statistics norm): absolute comparison
            assert float((p.grad - r.grad).norm()) <= 2e-2 * gmax, (k, float(p.grad.norm()), gmax)
            continue
        errs.append((rel_l2(p.grad, r.grad), k))
        if env_sd is not None and env_sd[k].grad is not None:
            env.append(rel_l2(env_sd[k].grad.float(), r.grad))
    errs.sort()
    assert errs
    med, worst = statistics.median(e for e, _ in errs), errs[-1]
    if mode == "linear":
        assert worst[0] <= LINEAR_TOL, (med, errs[-3:])
    else:
        # activation-flip envelope of a bf16 forward, measured on torch's own bf16 autocast of the oracle
        med_tol = max(MEDIAN_TOL, 1.25 * statistics.median(env)) if env else MEDIAN_TOL
        worst_tol = max(GRAD_FLIP_TOL, 1.25 * max(env)) if env else GRAD_FLIP_TOL
        assert med <= med_tol and worst[0] <= worst_tol, (med, med_tol, worst_tol, errs[-3:])


@pytest.mark.parametrize("mode", ["linear", "real"])
@pytest.mark.parametrize("norm", ["batch", "instance"])
def test_encoder_decoder_forward_backward(mode, norm, monkeypatch):
    from cycle_depth_estimation_b200 import encoder_decoder as E
    nc = 22
    enc, sd_e = _load(E._UNetEncoder(3, ngf=16, norm=norm), 11, OE.ENC_SLOPE, mode)
    dec, sd_d = _load(E._UNetDecoder(nc, ngf=16, norm=norm), 12, OE.DEC_SLOPE, mode)
    x0 = seeded_image(2, 3, 96, 128, seed=61)
    gout = seeded_image(2, nc, 96, 128, seed=62)
    g3 = seeded_image(2, nc, 24, 32, seed=63) * 0.5

    x = x0.clone().requires_grad_(True)
    feats = enc(x)
    outs = dec(feats)
    assert outs[0] is feats[3] and len(outs) == 5
    # the loss of models/seg_cycle.py:95-100 reads output[-1]; output3 gets a second seed so that a gradient arriving
    # at an intermediate output block is covered too
    ((outs[-1] * gout).sum() + (outs[2] * g3).sum()).backward()

    ref_e, ref_d = _leafs(sd_e), _leafs(sd_d)
    xr = x0.clone().requires_grad_(True)
    with true_fp32():
        rfeats = OE.unet_encoder(ref_e, xr)
        routs = OE.unet_decoder(ref_d, rfeats)
        ((routs[-1] * gout).sum() + (routs[2] * g3).sum()).backward()

    # what bf16 storage costs on this network whatever the kernels: torch's own bf16 autocast of the oracle against
    # its fp32 evaluation (the centre runs BatchNorm over 2x6x8 = 96 values per channel and amplifies rounding);
    # the B200 path must stay within max(2e-2, 1.25 x that envelope)
    env_e, env_d = _leafs(sd_e), _leafs(sd_d)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        efeats = OE.unet_encoder(env_e, x0)
        eouts = OE.unet_decoder(env_d, efeats)
        ((eouts[-1].float() * gout).sum() + (eouts[2].float() * g3).sum()).backward()
    for i, (a, b, e) in enumerate(zip(feats, rfeats, efeats)):
        assert a.shape == b.shape and a.dtype == torch.float32
        tol = max(TOL_BF16, 1.25 * rel_l2(e.detach().float(), b))
        assert rel_l2(a, b) <= tol, ("feat", i, rel_l2(a, b), tol)
    assert rel_l2(feats[0], rfeats[0]) <= TOL_BF16 and rel_l2(feats[1], rfeats[1]) <= TOL_BF16
    for i, (a, b, e) in enumerate(zip(outs[1:], routs[1:], eouts[1:])):
        assert a.shape == b.shape
        tol = max(TOL_BF16, 1.25 * rel_l2(e.detach().float(), b))
        assert rel_l2(a, b) <= tol, ("out", i, rel_l2(a, b), tol)
    tol_x = LINEAR_TOL if mode == "linear" else GRAD_FLIP_TOL
    assert rel_l2(x.grad, xr.grad) <= tol_x, rel_l2(x.grad, xr.grad)
    s_enc, s_dec = _slope_scales(sd_e, sd_d, x0, lambda o: (o[-1] * gout).sum() + (o[2] * g3).sum(), monkeypatch)
    _compare_grads(dec, ref_d, mode, OE.DEC_SLOPE, s_dec, env_d)
    _compare_grads(enc, ref_e, mode, OE.ENC_SLOPE, s_enc, env_e)
    if norm == "batch":     # running statistics follow torch's update
        after = enc.state_dict()
        for k in ('conv1.2.running_mean', 'conv3.model.4.running_var', 'center.2.norm1.running_mean'):
            assert rel_l2(after[k], ref_e[k], floor=1e-3) <= TOL_BF16, k
        assert int(after['conv1.2.num_batches_tracked']) == 1


def test_decoder_eval_mode_and_frozen_encoder():
    """eval(): BatchNorm uses the running statistics; a frozen encoder (requires_grad False, models/seg_cycle.py:160)
    gets no parameter gradients while the decoder still does."""
    from cycle_depth_estimation_b200 import encoder_decoder as E
    enc, sd_e = _load(E._UNetEncoder(3, ngf=16), 31, OE.ENC_SLOPE, "real")
    dec, sd_d = _load(E._UNetDecoder(6, ngf=16), 32, OE.DEC_SLOPE, "real")
    x = seeded_image(1, 3, 96, 96, seed=64)
    enc.eval()
    dec.eval()
    with torch.no_grad():
        outs = dec(enc(x))
        with true_fp32():
            routs = OE.unet_decoder(sd_d, OE.unet_encoder(sd_e, x, training=False), training=False)
    for a, b in zip(outs[1:], routs[1:]):
        assert rel_l2(a, b) <= 1.5 * TOL_BF16, rel_l2(a, b)
    enc.train()
    dec.train()
    for p in enc.parameters():
        p.requires_grad_(False)
    outs = dec(enc(x))
    outs[-1].sum().backward()
    assert all(p.grad is None for p in enc.parameters())
    assert all(p.grad is not None for p in dec.parameters())


def test_full_width_forward():
    """ngf=64, 28 classes, 256x256 (the SegCycle configuration, models/seg_cycle.py:48-51)."""
    from cycle_depth_estimation_b200 import encoder_decoder as E
    enc, sd_e = _load(E._UNetEncoder(3), 41, OE.ENC_SLOPE, "real")
    dec, sd_d = _load(E._UNetDecoder(28), 42, OE.DEC_SLOPE, "real")
    x = seeded_image(2, 3, 256, 256, seed=65)
    with torch.no_grad():
        outs = dec(enc(x))
        with true_fp32():
            routs = OE.unet_decoder(sd_d, OE.unet_encoder(sd_e, x))
        with torch.autocast('cuda', dtype=torch.bfloat16):
            eouts = OE.unet_decoder(sd_d, OE.unet_encoder(sd_e, x))
    assert tuple(outs[-1].shape) == (2, 28, 256, 256)
    for a, b, e in zip(outs[1:], routs[1:], eouts[1:]):
        tol = max(TOL_BF16, 1.25 * rel_l2(e.float(), b))
        assert rel_l2(a, b) <= tol, (rel_l2(a, b), tol)
