"""U-Net generator (models/networks.py:243-316) and the pix2pix step (models/pix2pix_model.py:66-111) on
the graph engine vs the fp32 oracle.  Tolerances as in test_networks_gpu.py: forward / losses <= 2e-2
relative L2; gradients inside the activation-flip envelope of a bf16 forward (GRAD_FLIP_TOL) and <= 2e-2 for
the same topology with the dropout-free, branch-free wiring check."""
import argparse

import pytest
import torch

from helpers import TOL_BF16, leaf_state, quiet, rel_l2, seeded_image, true_fp32
from oracle import networks_oracle as O

pytestmark = pytest.mark.gpu

GRAD_FLIP_TOL = 0.25


def _build_unet(num_downs, ngf=64, norm='batch', dropout=False):
    from cycle_depth_estimation_b200 import networks as N
    torch.manual_seed(3)
    with quiet():
        net = N.define_G(3, 3, ngf, 'unet_256' if num_downs == 8 else 'unet_128', norm, dropout, 'normal', 0.02,
                         ['cuda'])
    return net


@pytest.mark.parametrize("num_downs,size,batch", [(7, 128, 2), (8, 256, 2)])
def test_unet_forward(num_downs, size, batch):
    net = _build_unet(num_downs)
    x = seeded_image(batch, 3, size, size)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        got = net(x)
    with true_fp32(), torch.no_grad():
        ref = O.unet_generator(sd, x, num_downs, 'batch')
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert rel_l2(got, ref) <= TOL_BF16, rel_l2(got, ref)
    # running statistics of every BatchNorm were updated like torch does
    for k, v in net.state_dict().items():
        if 'running_' in k:
            assert rel_l2(v, sd[k], floor=1e-3) <= TOL_BF16, (k, rel_l2(v, sd[k], floor=1e-3))
        if k.endswith('num_batches_tracked'):
            assert int(v) == 1


def test_unet_backward():
    num_downs = 7
    net = _build_unet(num_downs)
    x0 = seeded_image(2, 3, 128, 128)
    gout = seeded_image(2, 3, 128, 128, seed=7)
    sd = leaf_state(net)
    x = x0.clone().requires_grad_(True)
    out = net(x)
    (out * gout).sum().backward()
    xr = x0.clone().requires_grad_(True)
    with true_fp32():
        ref = O.unet_generator(sd, xr, num_downs, 'batch')
        (ref * gout).sum().backward()
    assert rel_l2(out, ref) <= TOL_BF16, rel_l2(out, ref)
    assert rel_l2(x.grad, xr.grad) <= GRAD_FLIP_TOL, rel_l2(x.grad, xr.grad)
    named = dict(net.named_parameters())
    worst = ("", 0.0)
    for k, r in sd.items():
        if not r.requires_grad or r.grad is None:
            continue
        got = named[k].grad
        assert got is not None, k
        err = rel_l2(got, r.grad, floor=1e-6)
        worst = max(worst, (k, err), key=lambda t: t[1])
    assert worst[1] <= GRAD_FLIP_TOL, worst


def test_unet_state_dict_matches_reference_layout():
    net = _build_unet(8)
    keys = list(net.state_dict().keys())
    assert keys[0] == 'model.model.0.weight' and 'model.model.1.model.3.model.3.model.3.model.3.model.3.model.3.model.1.weight' in keys
    assert sum(p.numel() for p in net.parameters()) == 54413955


def test_unet_dropout_is_active_in_training_only():
    net = _build_unet(7, dropout=True)
    x = seeded_image(2, 3, 128, 128)
    with torch.no_grad():
        a, b = net(x), net(x)
        assert not torch.equal(a, b)
        net.eval()
        c, d = net(x), net(x)
        assert torch.equal(c, d)


def _opt():
    return argparse.Namespace(input_nc=3, output_nc=3, ngf=64, ndf=64, netG='unet_128', netD='basic', n_layers_D=3,
                              norm='batch', no_dropout=True, init_type='normal', init_gain=0.02, no_lsgan=True,
                              pool_size=0, lr=2e-4, beta1=0.5, lambda_L1=100.0, isTrain=True, device='cuda',
                              direction='AtoB')


def test_pix2pix_step_losses_and_updates():
    from cycle_depth_estimation_b200.pix2pix_model import Pix2PixModel
    torch.manual_seed(0)
    model = Pix2PixModel()
    with quiet():
        model.initialize(_opt())
    oracle = O.Pix2PixStepOracle(model.netG.state_dict(), model.netD.state_dict(), num_downs=7)
    a, b = seeded_image(4, 3, 128, 128, seed=21), seeded_image(4, 3, 128, 128, seed=22)
    before = {k: v.clone() for k, v in model.netG.state_dict().items()}
    model.set_input({'A': a, 'B': b, 'A_paths': None})
    model.optimize_parameters()
    got = model.get_current_losses()
    with true_fp32():
        ref = oracle.step(a, b)
    for k in ('G_GAN', 'G_L1', 'D_real', 'D_fake'):
        assert abs(got[k] - ref[k]) <= 3e-2 * max(abs(ref[k]), 1e-3), (k, got[k], ref[k])
    # Adam's first step moves every weight by lr * sign(grad): compare the update direction where the
    # reference gradient is not tiny
    after = model.netG.state_dict()
    k = 'model.model.0.weight'
    upd, upd_ref = after[k] - before[k], oracle.G[k].detach() - before[k]
    agree = float((torch.sign(upd) == torch.sign(upd_ref)).float().mean())
    assert agree > 0.9, agree
