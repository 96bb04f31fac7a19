"""Why CycleGANModel may batch the passes that share a network (cycle_gan_model.CycleGANModel._batched): with
InstanceNorm every sample is normalised on its own, so a pass over the concatenated batch returns, sample for sample,
what the separate passes of models/cycle_gan_model.py:80-99,111-137 return, and the mean-reduced losses over the two
equal halves are the two separate losses.  Checked here on the fp32 oracle networks (CPU); the GPU step tests compare
the batched B200 step with the oracle's UNBATCHED step."""
import argparse

import torch

from helpers import quiet, rel_l2
from oracle import networks_oracle as O


def _sd(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def test_instance_norm_networks_commute_with_batch_concatenation():
    from cycle_depth_estimation_b200 import networks as N
    torch.manual_seed(0)
    with quiet():
        g = N.define_G(3, 3, 8, 'resnet_6blocks', 'instance', False, 'normal', 0.02, ['cpu'])
        d = N.define_D(3, 8, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu'])
    sg, sdd = _sd(g), _sd(d)
    gen = torch.Generator().manual_seed(3)
    a, b, c = (torch.rand((2, 3, 32, 32), generator=gen) * 2 - 1 for _ in range(3))
    with torch.no_grad():
        joint = O.resnet_generator(sg, torch.cat([a, b, c], 0), 6)
        parts = [O.resnet_generator(sg, x, 6) for x in (a, b, c)]
        assert rel_l2(joint, torch.cat(parts, 0)) < 1e-6
        real, fake = torch.rand((2, 3, 64, 64), generator=gen), torch.rand((2, 3, 64, 64), generator=gen)
        pred = O.nlayer_discriminator(sdd, torch.cat([real, fake], 0))
        sep = (O.gan_loss(O.nlayer_discriminator(sdd, real), True) + O.gan_loss(O.nlayer_discriminator(sdd, fake), False)) * 0.5
        bat = (O.gan_loss(pred[:2], True) + O.gan_loss(pred[2:], False)) * 0.5
        assert abs(float(sep) - float(bat)) < 1e-6 * max(1.0, abs(float(sep)))


def test_batch_norm_networks_do_not_commute_and_are_not_batched():
    from cycle_depth_estimation_b200 import networks as N
    from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel
    torch.manual_seed(0)
    with quiet():
        d = N.define_D(3, 8, 'basic', 3, 'batch', False, 'normal', 0.02, ['cpu'])
    sdd = _sd(d)
    gen = torch.Generator().manual_seed(4)
    real, fake = torch.rand((2, 3, 64, 64), generator=gen), torch.rand((2, 3, 64, 64), generator=gen) * 0.2
    with torch.no_grad():
        joint = O.nlayer_discriminator({k: v.clone() for k, v in sdd.items()}, torch.cat([real, fake], 0), 'batch')
        alone = O.nlayer_discriminator({k: v.clone() for k, v in sdd.items()}, real, 'batch')
    assert rel_l2(joint[:2], alone) > 1e-3          # batch statistics couple the samples
    m = CycleGANModel()
    m.opt = argparse.Namespace(norm='batch')
    assert not m._batched()
    m.opt = argparse.Namespace(norm='instance')
    assert m._batched()
    m.opt = argparse.Namespace(norm='instance', batch_passes=False)
    assert not m._batched()
