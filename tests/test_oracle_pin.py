"""Pins the oracle restatement (oracle/networks_oracle.py):
  (a) against the committed fixtures the REFERENCE's own modules produced (oracle/make_golden.py), and
  (b) against the reference modules themselves where /root/reference exists (build container only).
CPU only. Tolerances: same fp32 torch ops in a different composition -> 1e-5 relative; ImagePool ids and
threshold fractions exact."""
import os
import random

import numpy as np
import pytest
import torch

from helpers import quiet, rel_l2
from oracle import make_golden as MG
from oracle import networks_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def nets():
    return torch.load(os.path.join(GOLD, "networks.pt"), weights_only=False)


def test_resnet_generator_forward_and_grads(nets):
    fx = nets['resnet']
    sd = {k: v.clone().requires_grad_(True) for k, v in fx['sd'].items()}
    x = fx['x'].clone().requires_grad_(True)
    out = O.resnet_generator(sd, x, n_blocks=6)
    assert rel_l2(out, fx['out']) < 1e-5
    g = torch.Generator().manual_seed(5)
    (out * (torch.rand((2, 3, 32, 32), generator=g) * 2 - 1)).sum().backward()
    assert rel_l2(x.grad, fx['gx']) < 1e-4
    for k, ref in fx['gw'].items():
        if float(ref.norm()) < 1e-6:
            assert float(sd[k].grad.abs().max()) < 1e-5, k
        else:
            assert rel_l2(sd[k].grad, ref) < 1e-4, k


@pytest.mark.parametrize("norm", ["instance", "batch"])
def test_nlayer_discriminator(nets, norm):
    fx = nets['nlayer_' + norm]
    sd = {k: v.clone() for k, v in fx['sd'].items()}
    out = O.nlayer_discriminator(sd, fx['x'], norm=norm, use_sigmoid=(norm == 'batch'))
    assert rel_l2(out, fx['out']) < 1e-5
    if norm == 'batch':  # running statistics are updated like nn.BatchNorm2d does
        for k in sd:
            if 'running' in k:
                assert rel_l2(sd[k], fx['sd_after'][k]) < 1e-5, k


def test_unet_generator_including_inplace_skip_quirk(nets):
    fx = nets['unet']
    out = O.unet_generator({k: v.clone() for k, v in fx['sd'].items()}, fx['x'], num_downs=7, norm='batch')
    assert rel_l2(out, fx['out']) < 1e-5


def test_pixel_discriminator_and_gan_loss(nets):
    fx = nets['pixel']
    assert rel_l2(O.pixel_discriminator(fx['sd'], fx['x']), fx['out']) < 1e-5
    gl = nets['gan_loss']
    p = gl['pred']
    assert abs(float(O.gan_loss(p, True)) - gl['lsgan_real']) < 1e-6
    assert abs(float(O.gan_loss(p, False)) - gl['lsgan_fake']) < 1e-6
    assert abs(float(O.gan_loss(torch.sigmoid(p), True, use_lsgan=False)) - gl['bce_real']) < 1e-6
    assert abs(float(O.gan_loss(torch.sigmoid(p), False, use_lsgan=False)) - gl['bce_fake']) < 1e-6


def test_image_pool_returned_ids():
    fx = torch.load(os.path.join(GOLD, "image_pool.pt"), weights_only=False)
    random.seed(fx['seed'])
    pool = O.ImagePoolOracle(fx['pool_size'])
    nxt = 0
    for want in fx['returned']:
        b = len(want)
        batch = torch.stack([torch.full((1, 2, 2), float(nxt + i)) for i in range(b)])
        nxt += b
        got = [int(v) for v in pool.query(batch)[:, 0, 0, 0]]
        assert got == want


def test_image_pool_product_class_matches_golden_on_cpu():
    """The product ImagePool is host logic + tensor copies, so its decisions can be pinned on CPU too."""
    from cycle_depth_estimation_b200.image_pool import ImagePool
    fx = torch.load(os.path.join(GOLD, "image_pool.pt"), weights_only=False)
    random.seed(fx['seed'])
    pool = ImagePool(fx['pool_size'])
    nxt = 0
    for want in fx['returned']:
        b = len(want)
        batch = torch.stack([torch.full((1, 2, 2), float(nxt + i)) for i in range(b)])
        nxt += b
        assert [int(v) for v in pool.query(batch)[:, 0, 0, 0]] == want


def _pair(h, w, seed):
    rng = np.random.default_rng(seed)
    gt = rng.integers(0, 80, (h, w), dtype=np.uint8)
    gt[rng.random((h, w)) < 0.3] = 0
    return gt, rng.integers(0, 256, (h, w), dtype=np.uint8)


def test_depth_metrics_rows():
    fx = torch.load(os.path.join(GOLD, "metrics.pt"), weights_only=False)
    for (h, w, seed), want in zip(fx['pairs'], fx['rows']):
        gt, pred = _pair(h, w, seed)
        p = np.clip(pred / 255 * 80, 1, 50)
        mask = np.logical_and(gt > 1, gt < 50)
        got = O.compute_errors(gt[mask], p[mask])
        assert got[4:] == tuple(want[4:])                       # threshold fractions: exact
        assert np.allclose(got[:4], want[:4], rtol=1e-12, atol=0)
        means, per = O.eval_metric_arrays([gt], [pred])
        assert np.allclose(per[0], np.asarray(want, np.float32), rtol=1e-6)


# ---- live reference (build container only) ---------------------------------------------------------
live = pytest.mark.skipif(not MG.available(), reason="/root/reference not present")


@live
def test_live_reference_cyclegan_networks_at_full_width():
    N = MG.load_ref("ref_networks_live", "models/networks.py")
    torch.manual_seed(0)
    with quiet():
        g = N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu'])
        d = N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu'])
    x = MG.image(1, 3, 64, 64, 3)
    with torch.no_grad():
        assert rel_l2(O.resnet_generator(g.state_dict(), x, 9), g(x)) < 1e-5
        assert rel_l2(O.nlayer_discriminator(d.state_dict(), x), d(x)) < 1e-5


@live
def test_live_reference_state_dict_layout_matches_product_modules():
    N = MG.load_ref("ref_networks_live2", "models/networks.py")
    from cycle_depth_estimation_b200 import networks as M
    cases = [(lambda mod: mod.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu'])),
             (lambda mod: mod.define_G(3, 3, 64, 'resnet_6blocks', 'batch', True, 'normal', 0.02, ['cpu'])),
             (lambda mod: mod.define_G(3, 3, 64, 'unet_256', 'batch', True, 'normal', 0.02, ['cpu'])),
             (lambda mod: mod.define_G(3, 3, 16, 'unet_128', 'instance', False, 'normal', 0.02, ['cpu'])),
             (lambda mod: mod.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu'])),
             (lambda mod: mod.define_D(6, 64, 'basic', 3, 'batch', True, 'normal', 0.02, ['cpu'])),
             (lambda mod: mod.define_D(3, 64, 'pixel', 3, 'batch', False, 'normal', 0.02, ['cpu']))]
    for make in cases:
        torch.manual_seed(0)
        with quiet():
            a = make(N)
        torch.manual_seed(0)
        with quiet():
            b = make(M)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        for k in sa:
            assert sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype, k
            assert torch.equal(sa[k], sb[k]), k          # same init stream -> identical values
        b.load_state_dict(sa, strict=True)
    for bad in (lambda: M.define_G(3, 3, 64, '3blocks', 'instance', False, 'normal', 0.02, ['cpu']),
                lambda: M.define_D(3, 64, 'n_layers', 3, 'instance', False, 'normal', 0.02, ['cpu']),
                lambda: M.get_norm_layer('group')):
        with pytest.raises(NotImplementedError):
            bad()
    with pytest.raises(IndexError):
        with quiet():
            M.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, [])


@live
def test_live_reference_eval_metric_semantics():
    E = MG.load_ref("ref_my_eval_live", "new_multi/my_eval.py")
    gt, pred = _pair(120, 160, 5)
    p = np.clip(pred / 255 * 80, 1, 50)
    mask = np.logical_and(gt > 1, gt < 50)
    with quiet():
        want = E.compute_errors(gt[mask], p[mask])
    got = O.compute_errors(gt[mask], p[mask])
    assert got[4:] == want[4:] and np.allclose(got[:4], want[:4], rtol=1e-12, atol=0)
