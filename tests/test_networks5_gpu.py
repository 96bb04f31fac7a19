"""The seg/depth networks (new_multi/networks5_ds.py: G_1, General_net, R_dep, _Discriminator) on the graph
engine vs the fp32 oracle (oracle/networks5_oracle.py, pinned to the reference's own classes) on identical
name-keyed synthetic weights.  Tolerances: forward <= 2e-2 relative L2 (bf16 path); gradients inside the
activation-flip envelope of a bf16 forward (see test_networks_gpu.py), with the median parameter-gradient
error required to be far below it so that a wiring error (O(1)) cannot hide."""
import statistics

import pytest
import torch

from helpers import TOL_BF16, quiet, rel_l2, seeded_image, true_fp32
from oracle import networks5_oracle as O5

pytestmark = pytest.mark.gpu

GRAD_FLIP_TOL = 0.35     # envelope with ~20 activation layers (branch flips of a bf16 forward)
MEDIAN_TOL = 0.2
LINEAR_TOL = 4e-2        # same wiring with every ReLU / LeakyReLU / PReLU made the identity: no branch can flip


@pytest.fixture(params=["linear", "real"])
def mode(request, monkeypatch):
    """'linear' removes every activation branch from BOTH implementations (BatchNorm keeps the networks
    non-trivial), so the hand-written backward wiring — concat-slice gradients, fp32 dense-block accumulators,
    attention gates, re-sampling, residual adds — must match fp32 torch to LINEAR_TOL end to end."""
    if request.param == "linear":
        import torch.nn.functional as F
        from cycle_depth_estimation_b200 import networks5_ds as N
        monkeypatch.setattr(N, "ACT_RELU", N.ACT_NONE)
        monkeypatch.setattr(N, "ACT_LEAKY", N.ACT_NONE)
        monkeypatch.setattr(F, "relu", lambda x, *a, **k: x)
        monkeypatch.setattr(F, "leaky_relu", lambda x, *a, **k: x)
    return request.param


def _load(net, seed, mode="real"):
    sd = O5.synth_state_dict(net.state_dict(), seed)
    if 'model.10.weight' in sd and 'model.1.weight' in sd:
        sd['model.1.weight'] = sd['model.10.weight']
        if mode == "linear":
            for k in ('model.1.weight', 'model.4.weight', 'model.7.weight', 'model.10.weight'):
                sd[k] = torch.ones_like(sd[k])
    net.load_state_dict(sd, strict=True)
    return net.cuda().train(), {k: v.cuda() for k, v in sd.items()}


def _grad_report(net, ref_sd):
    named = dict(net.named_parameters())
    errs = []
    gmax = max(float(r.grad.norm()) for r in ref_sd.values() if r.requires_grad and r.grad is not None)
    for k, r in ref_sd.items():
        if not r.requires_grad:
            continue
        if k not in named:          # second key of a shared parameter
            continue
        got = named[k].grad
        if r.grad is None:
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        assert got is not None, "missing gradient for " + k
        if float(r.grad.norm()) < 1e-3 * gmax:
            # mathematically (near-)zero gradient, e.g. a BatchNorm bias feeding another batch-statistics
            # BatchNorm: rounding noise in both implementations -> absolute comparison against the scale
            assert float(got.norm()) <= 2e-2 * gmax, (k, float(got.norm()), gmax)
            continue
        errs.append((rel_l2(got, r.grad), k))
    errs.sort()
    return errs


def _check_grads(net, ref_sd, mode="real", env_sd=None):
    """env_sd: the oracle's parameters after a backward pass under torch's bf16 autocast — the gradient
    envelope of bf16 storage on this network (used for the 82-layer trunk, see _bf16_envelope)."""
    errs = _grad_report(net, ref_sd)
    assert errs, "no gradients compared"
    worst = errs[-1]
    med = statistics.median(e for e, _ in errs)
    if mode == "linear":
        assert worst[0] <= LINEAR_TOL, (med, worst)
        return
    med_tol, worst_tol = MEDIAN_TOL, GRAD_FLIP_TOL
    if env_sd is not None:
        env = sorted(rel_l2(env_sd[k].grad, r.grad) for k, r in ref_sd.items()
                     if r.requires_grad and r.grad is not None and env_sd[k].grad is not None
                     and float(r.grad.norm()) > 0)
        med_tol = max(med_tol, 1.25 * statistics.median(env))
        worst_tol = max(worst_tol, 1.25 * env[-1])
    assert med <= med_tol, (med, med_tol, worst)
    assert worst[0] <= worst_tol, (worst, worst_tol)


def _bf16_envelope(fn, ref_outs):
    """Errors of torch's OWN bf16 autocast of the oracle against its fp32 evaluation: what bf16 storage costs
    on this network whatever the kernels. The 82-layer DenseNet trunk amplifies bf16 rounding far beyond 2e-2
    (13 % at the head with these synthetic weights, DESIGN.md 'tolerances'); the B200 path is required to stay
    within 1.25x of that envelope (or 2e-2, whichever is larger)."""
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
        outs = fn()
    return [max(TOL_BF16, 1.25 * rel_l2(a.float(), r)) for a, r in zip(outs, ref_outs)]


def _gtol(mode):
    return LINEAR_TOL if mode == "linear" else GRAD_FLIP_TOL


def test_g1_forward_backward(mode):
    from cycle_depth_estimation_b200 import networks5_ds as N
    net, sd = _load(N.G_1(), 1, mode)
    x0 = seeded_image(2, 3, 64, 128, seed=41)
    gout = seeded_image(2, 64, 32, 64, seed=42)
    x = x0.clone().requires_grad_(True)
    out = net(x)
    (out * gout).sum().backward()
    ref_sd = O5.leaf_params(sd)
    xr = x0.clone().requires_grad_(True)
    with true_fp32():
        ref = O5.g_1(ref_sd, xr)
        (ref * gout).sum().backward()
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel_l2(out, ref) <= TOL_BF16, rel_l2(out, ref)
    assert rel_l2(x.grad, xr.grad) <= _gtol(mode), rel_l2(x.grad, xr.grad)
    _check_grads(net, ref_sd, mode)
    # BatchNorm running statistics follow torch's update
    after = net.state_dict()
    for k in ('features.norm0.running_mean', 'features.denseblock1.denselayer6.norm1.running_var',
              'model.6.conv1_block.2.running_mean'):
        assert rel_l2(after[k], ref_sd[k], floor=1e-3) <= TOL_BF16, k


@pytest.mark.parametrize("kind", ["S", "R"])
def test_general_net_forward_backward(kind, mode):
    from cycle_depth_estimation_b200 import networks5_ds as N
    net, sd = _load(N.General_net(), 2, mode)
    x0 = seeded_image(2, 64, 32, 64, seed=43) if kind == 'S' else seeded_image(2, 3, 64, 128, seed=44)
    x = x0.clone().requires_grad_(True)
    head, feats = net(x, kind)
    gout = seeded_image(*head.shape, seed=45)
    (head * gout).sum().backward()
    ref_sd = O5.leaf_params(sd)
    xr = x0.clone().requires_grad_(True)
    with true_fp32():
        rhead, rfeats = O5.general_net(ref_sd, xr, kind)
        (rhead * gout).sum().backward()
    def auto():
        h, fs = O5.general_net({k: v.detach().clone() for k, v in ref_sd.items()}, x0, kind)
        return [h] + fs
    tol = _bf16_envelope(auto, [rhead] + rfeats)
    assert rel_l2(head, rhead) <= tol[0], (rel_l2(head, rhead), tol[0])
    assert len(feats) == 4
    for f, r, t in zip(feats, rfeats, tol[1:]):
        assert f.shape == r.shape and not f.requires_grad
        assert rel_l2(f, r) <= t, (rel_l2(f, r), t)
    assert rel_l2(feats[0], rfeats[0]) <= TOL_BF16
    env_sd, xtol = None, _gtol(mode)
    if mode == "real":
        env_sd = O5.leaf_params(sd)
        xa = x0.clone().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            ahead, _ = O5.general_net(env_sd, xa, kind)
        (ahead.float() * gout).sum().backward()
        xtol = max(xtol, 1.25 * rel_l2(xa.grad, xr.grad))
    assert rel_l2(x.grad, xr.grad) <= xtol, (rel_l2(x.grad, xr.grad), xtol)
    _check_grads(net, ref_sd, mode, env_sd)


def _rdep_inputs():
    s = [None, seeded_image(2, 512, 16, 32, seed=51), seeded_image(2, 1280, 8, 16, seed=52),
         seeded_image(2, 1664, 4, 8, seed=53)]
    s[0] = seeded_image(2, 256, 32, 64, seed=50)
    return s, seeded_image(2, 1024, 4, 8, seed=54)


def test_r_dep_forward_backward(mode):
    from cycle_depth_estimation_b200 import networks5_ds as N
    net, sd = _load(N.R_dep(), 3, mode)
    s, d0 = _rdep_inputs()
    d = d0.clone().requires_grad_(True)
    feats, seg, (dep4, dep1) = net(s, d)
    ref_sd = O5.leaf_params(sd)
    dr = d0.clone().requires_grad_(True)
    with true_fp32():
        rfeats, rseg, (rdep4, rdep1) = O5.r_dep(ref_sd, s, dr)
    pairs = list(zip(feats, rfeats)) + [(seg, rseg), (dep1, rdep1)] + list(zip(dep4, rdep4))

    def auto():
        f, sg, (d4, d1) = O5.r_dep({k: v.detach().clone() for k, v in ref_sd.items()}, s, d0)
        return list(f) + [sg, d1] + list(d4)
    tol = _bf16_envelope(auto, [r for _, r in pairs])
    for (got, ref), t in zip(pairs, tol):
        assert got.shape == ref.shape and got.dtype == torch.float32
        assert rel_l2(got, ref) <= t, (tuple(got.shape), rel_l2(got, ref), t)
    # one scalar objective touching every output
    gs = [seeded_image(*ref.shape, seed=60 + i) for i, (_, ref) in enumerate(pairs)]
    sum((g * o).sum() for g, (o, _) in zip(gs, pairs)).backward()
    with true_fp32():
        sum((g * r).sum() for g, (_, r) in zip(gs, pairs)).backward()
    assert rel_l2(d.grad, dr.grad) <= _gtol(mode), rel_l2(d.grad, dr.grad)
    _check_grads(net, ref_sd, mode)
    # modules that R_dep.forward never calls get no gradient
    assert net.up0.deconv.weight.grad is None and net.dep_out.weight.grad is None


def test_feature_discriminator_shared_prelu(mode):
    from cycle_depth_estimation_b200 import networks5_ds as N
    net, sd = _load(N._Discriminator(input_nc=128), 4, mode)
    assert net.model[1] is net.model[10]
    x0 = seeded_image(2, 128, 32, 64, seed=70)
    x = x0.clone().requires_grad_(True)
    out = net(x)
    gout = seeded_image(*out.shape, seed=71)
    (out * gout).sum().backward()
    ref_sd = O5.leaf_params(sd)
    ref_sd['model.10.weight'] = ref_sd['model.1.weight']
    xr = x0.clone().requires_grad_(True)
    with true_fp32():
        ref = O5.discriminator(ref_sd, xr)
        (ref * gout).sum().backward()
    assert rel_l2(out, ref) <= TOL_BF16, rel_l2(out, ref)
    assert rel_l2(x.grad, xr.grad) <= _gtol(mode)
    _check_grads(net, ref_sd, mode)
    assert rel_l2(net.model[1].weight.grad, ref_sd['model.1.weight'].grad) <= 0.1


def test_init_net_keeps_the_dataparallel_key_prefix():
    from cycle_depth_estimation_b200 import networks5_ds as N
    with quiet():
        net = N.init_net(N._Discriminator(input_nc=64))
    assert all(k.startswith('module.') for k in net.state_dict())
    assert net(seeded_image(1, 64, 32, 32)).shape == (1, 1, 2, 2)


def test_losses_of_the_step():
    from cycle_depth_estimation_b200 import losses, networks5_ds as N
    import torch.nn.functional as F
    logits = seeded_image(2, 28, 16, 24, seed=80).requires_grad_(True)
    g = torch.Generator().manual_seed(81)
    labels = torch.randint(0, 28, (2, 16, 24), generator=g)
    labels[0, :3] = 255
    labels = labels.cuda()
    loss = losses.CrossEntropyLoss(ignore_index=255)(logits, labels)
    (loss * 2.0).backward()
    lr = logits.detach().clone().requires_grad_(True)
    ref = F.cross_entropy(lr, labels, ignore_index=255)
    (ref * 2.0).backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_l2(logits.grad, lr.grad) <= 1e-5
    t = seeded_image(2, 4, 16, 24, seed=82)
    t[t > 0.6] = 1.0
    t[t < -0.6] = -1.0
    o_m, z_m = N.get_masks(t)
    ro, rz = O5.get_masks(t)
    assert torch.equal(o_m, ro) and torch.equal(z_m, rz)
    x = torch.tanh(seeded_image(2, 1, 16, 24, seed=83)).requires_grad_(True)
    l2 = N.BCEDepLoss()(x, t, o_m, z_m)
    l2.backward()
    xr = x.detach().clone().requires_grad_(True)
    r2 = O5.bce_dep_loss(xr, t, ro, rz)
    r2.backward()
    assert abs(float(l2) - float(r2)) <= 1e-5 * abs(float(r2))
    assert rel_l2(x.grad, xr.grad) <= 1e-5
    # L1 with the reference's [B,1,H,W] vs [B,H,W] broadcasting
    a = seeded_image(3, 1, 8, 8, seed=84).requires_grad_(True)
    b = seeded_image(3, 8, 8, 1, seed=85)[..., 0]
    l3 = losses.L1Loss()(a, b)
    l3.backward()
    ar = a.detach().clone().requires_grad_(True)
    r3 = F.l1_loss(ar.expand(3, 3, 8, 8), b.expand(3, 3, 8, 8))
    r3.backward()
    assert abs(float(l3) - float(r3)) <= 1e-5 * abs(float(r3)) and rel_l2(a.grad, ar.grad) <= 1e-5
    assert float(N.GANLoss(use_lsgan=True)(x.detach(), True)) == pytest.approx(float(((x.detach() - 1) ** 2).mean()), rel=1e-5)
