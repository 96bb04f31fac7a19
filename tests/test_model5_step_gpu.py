"""The seg/depth training step (new_multi/model5.py:640-696) on the B200 networks vs the restated step of
oracle/networks5_oracle.py on identical initial weights and inputs.

Tolerances.  The segmentation / depth losses (G2, G1, RD_real, RD_syn, dep_ref) must agree within 3e-2 relative
in the first step and 1e-1 in the second (i.e. after all eight optimizer updates of the first).  The three
feature-discriminator losses are means over a handful of PatchGAN outputs of features that went through the
82-layer DenseNet trunk, where bf16 storage alone moves activations by > 10 % (test_networks5_gpu.py,
DESIGN.md 'tolerances'); measured here: R_dep's features differ from fp32 by 22-27 % (real) / 44-53 % (synthetic branch) both for this path
and for torch's OWN bf16 autocast of the oracle step.  Those losses (means over 4-150 values of O(1) outputs
of randomly initialised PatchGANs) are therefore noise-dominated in ANY bf16 evaluation and are only bounded
loosely: |ours - fp32| <= max(0.5 |fp32|, 3 |autocast - fp32|).  The discriminators themselves are pinned
tightly on identical inputs in test_networks5_gpu.py."""
import argparse

import pytest
import torch

from helpers import quiet, true_fp32
from oracle import networks5_oracle as O5

pytestmark = pytest.mark.gpu


def _inputs(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(s, generator=g) * 2 - 1
    seg_syn = torch.randint(0, 28, (b, 1, h, w), generator=g)
    seg_real = torch.randint(0, 28, (b, 1, h, w), generator=g)
    seg_real[torch.rand((b, 1, h, w), generator=g) < 0.02] = 255
    dls = r(b, 4, h, w)
    dls[dls > 0.9] = 1.0
    dls[dls < -0.9] = -1.0
    return {'img_real': r(b, 3, h, w), 'img_syn': r(b, 3, h, w), 'seg_l_real': seg_real, 'seg_l_syn': seg_syn,
            'dep_l_syn': r(b, 1, h, w), 'depth_l_s': dls}


def test_seg_depth_step_losses():
    from cycle_depth_estimation_b200.model5 import Seg_Depth
    torch.manual_seed(0)
    model = Seg_Depth()
    with quiet():
        model.initialize(argparse.Namespace(lr=2e-4, beta1=0.5, pool_size=50))
    strip = O5.strip_module_prefix
    sds = [strip(getattr(model, 'net_' + n).state_dict()) for n in ('G_1', 'G_2', 'R_D', 'FD1', 'FD2', 'FD3')]
    oracle = O5.SegDepthStepOracle(*sds)
    envelope = O5.SegDepthStepOracle(*sds)
    data = _inputs(2, 192, 256, 90)
    cu = {k: v.cuda() for k, v in data.items()}
    for step, tol in ((0, 3e-2), (1, 1e-1)):
        model.set_input(data, 'train')
        model.optimize_parameters('train')
        got = model.get_current_losses()
        with true_fp32():
            ref = oracle.step(cu['img_syn'], cu['img_real'], cu['seg_l_syn'].squeeze(1), cu['seg_l_real'].squeeze(1),
                              cu['dep_l_syn'].squeeze(1), cu['depth_l_s'])
        args = (cu['img_syn'], cu['img_real'], cu['seg_l_syn'].squeeze(1), cu['seg_l_real'].squeeze(1),
                cu['dep_l_syn'].squeeze(1), cu['depth_l_s'])
        with torch.autocast('cuda', dtype=torch.bfloat16):
            env = envelope.step(*args)
        print(step, {k: (round(got[k], 4), round(ref[k], 4), round(env[k], 4)) for k in ref if k in got})
        for k in ('G2', 'G1', 'RD_real', 'RD_syn', 'dep_ref'):
            assert abs(got[k] - ref[k]) <= tol * max(abs(ref[k]), 1e-3), (step, k, got[k], ref[k])
        for k in ('FD1', 'FD2', 'FD3'):
            bound = max(0.5 * abs(ref[k]), 3.0 * abs(env[k] - ref[k]))
            assert abs(got[k] - ref[k]) <= bound, (step, k, got[k], ref[k], env[k])
    assert model.syn_dep_ref.shape == (2, 192, 256) and model.real_dep_ref.shape == (2, 192, 256)
