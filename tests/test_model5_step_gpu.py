"""The seg/depth training step (new_multi/model5.py:640-696) on the B200 networks vs the restated step of
oracle/networks5_oracle.py on identical initial weights and inputs.

Tolerances.  The segmentation / depth losses (G2, G1, RD_real, RD_syn, dep_ref) must agree within 3e-2 relative
in the first step and 1e-1 in the second (i.e. after all eight optimizer updates of the first).

The three feature-discriminator losses are means over 4-150 outputs of randomly initialised PatchGANs applied
to R_dep features that went through the 82-layer DenseNet trunk at initialisation.  In that regime the network
is chaotic: R_dep's features differ from the fp32 oracle by 22-27 % (real) / 44-53 % (synthetic branch) for this
path AND for torch's own bf16 autocast of the oracle (measured, tools/debug_m5.py), and two runs of this path on
identical inputs differ from each other (atomic summation order of the statistics).  An end-to-end comparison
of those three numbers would compare noise, so they are checked by teacher forcing instead: the fp32 oracle
discriminator (initial weights) is evaluated on the features THIS path fed to its discriminators, and the losses
this path reported must match those within 3e-2.  The discriminators themselves are also pinned on identical
inputs in test_networks5_gpu.py."""
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
    sds = [{k: v.clone() for k, v in strip(getattr(model, 'net_' + n).state_dict()).items()}
           for n in ('G_1', 'G_2', 'R_D', 'FD1', 'FD2', 'FD3')]
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
            bound = tol * max(abs(ref[k]), 1e-3)
            if k == 'RD_real' and step > 0:
                # RD_real = CE + 0.2 x three feature-discriminator terms; after the first updates those terms are
                # the chaotic quantities described above (second step: 15.0 fp32 / 12.0 autocast / 11.6 here for
                # FD1), so the bf16-autocast envelope bounds this one
                bound = max(bound, 2.0 * abs(env[k] - ref[k]) + 0.05 * abs(ref[k]))
            assert abs(got[k] - ref[k]) <= bound, (step, k, got[k], ref[k], env[k])
        for i, k in enumerate(('FD1', 'FD2', 'FD3')):
            assert got[k] == got[k] and 0.0 < got[k] < 50.0, (step, k, got[k])
            if step == 0:
                with torch.no_grad(), true_fp32():
                    fd = {n: v.clone() for n, v in sds[3 + i].items()}
                    tf = (O5.gan_mse(O5.discriminator(fd, model.real_feats[i].detach()), True)
                          + O5.gan_mse(O5.discriminator(fd, model.syn_feats[i].detach()), False))
                assert abs(got[k] - float(tf)) <= 3e-2 * abs(float(tf)), (k, got[k], float(tf), ref[k], env[k])
    assert model.syn_dep_ref.shape == (2, 192, 256) and model.real_dep_ref.shape == (2, 192, 256)


def test_seg_depth_cuda_graph_step_follows_the_eager_step():
    """opt.cuda_graph: 3 eager warm-up steps, capture, replays. The stable losses must follow the eager run."""
    from cycle_depth_estimation_b200.model5 import Seg_Depth

    def run(graph):
        torch.manual_seed(0)
        model = Seg_Depth()
        with quiet():
            model.initialize(argparse.Namespace(lr=2e-4, beta1=0.5, pool_size=50, cuda_graph=graph))
        hist = []
        for step in range(6):
            model.set_input(_inputs(1, 192, 256, 300 + step), 'train')
            model.optimize_parameters('train')
            hist.append(model.get_current_losses())
        return hist, model

    eager, _ = run(False)
    graphed, model = run(True)
    assert model._step_graph.graph is not None and model._step_graph.launches > 1000
    for step, (e, g) in enumerate(zip(eager, graphed)):
        for k in ('G2', 'G1', 'RD_syn', 'dep_ref'):
            assert abs(e[k] - g[k]) <= 5e-2 * max(abs(e[k]), 1e-2), (step, k, e[k], g[k])
        # RD_real contains 0.2 x the three (chaotic, see the module docstring) feature-discriminator terms
        assert abs(e['RD_real'] - g['RD_real']) <= 0.2 * abs(e['RD_real']), (step, e['RD_real'], g['RD_real'])
