"""Pins oracle/networks5_oracle.py against fixtures the REFERENCE's own new_multi/networks5_ds.py classes
produced (oracle/make_golden.py -> tests/golden/networks5.pt; weights regenerated here from names + seed).
CPU only. Tolerance: the same fp32 torch ops in a different composition -> 2e-5 relative L2."""
import os

import pytest
import torch

from helpers import rel_l2
from oracle import networks5_oracle as O5

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLD, "networks5.pt"), weights_only=False)


@pytest.fixture(scope="module")
def nets():
    from cycle_depth_estimation_b200 import networks5_ds as N
    return N


def _sd(module, seed):
    return O5.synth_state_dict(module.state_dict(), seed)


def test_g1_and_general_net(fx, nets):
    sd1 = _sd(nets.G_1(), 1)
    ss = O5.g_1(sd1, fx['g1']['x'])
    assert rel_l2(ss, fx['g1']['out']) < TOL
    sd2 = _sd(nets.General_net(), 2)
    head, feats = O5.general_net({k: v.clone() for k, v in sd2.items()}, fx['g1']['out'], 'S')
    assert rel_l2(head, fx['g2_S']['head']) < TOL
    assert rel_l2(feats[3], fx['g2_S']['feat3']) < TOL
    for f, m, a in zip(feats, fx['g2_S']['feat_means'], fx['g2_S']['feat_abs']):
        assert abs(float(f.mean()) - m) < 1e-5 and abs(float(f.abs().mean()) - a) < 1e-5
        assert not f.requires_grad
    head_r, feats_r = O5.general_net({k: v.clone() for k, v in sd2.items()}, fx['g2_R']['x'], 'R')
    assert rel_l2(head_r, fx['g2_R']['head']) < TOL
    for f, a in zip(feats_r, fx['g2_R']['feat_abs']):
        assert abs(float(f.abs().mean()) - a) < 1e-5


def test_r_dep_and_discriminator(fx, nets):
    sd2 = _sd(nets.General_net(), 2)
    head, feats = O5.general_net(sd2, fx['g1']['out'], 'S')
    sd3 = _sd(nets.R_dep(), 3)
    (o0, o1, o2), seg, (dep4, dep1) = O5.r_dep(sd3, feats, head.detach())
    r = fx['rd']
    assert rel_l2(o0, r['out0']) < TOL and rel_l2(o1, r['out1']) < TOL
    assert rel_l2(o2[:, :, ::2, ::2], r['out2']) < TOL
    assert rel_l2(seg[:, :, ::2, ::2], r['seg']) < TOL
    for a, b in zip(dep4, r['dep4']):
        assert rel_l2(a, b) < TOL
    assert rel_l2(dep1, r['dep1']) < TOL
    sd4 = _sd(nets._Discriminator(input_nc=128), 4)
    sd4['model.1.weight'] = sd4['model.10.weight']   # ONE shared nn.PReLU: load_state_dict leaves the later key's value
    assert rel_l2(O5.discriminator(sd4, fx['fd']['x']), fx['fd']['out']) < TOL


def test_masks_and_bcedep(fx):
    b = fx['bcedep']
    o_m, z_m = O5.get_masks(b['t'])
    assert torch.equal(o_m, b['o_m']) and torch.equal(z_m, b['z_m'])
    assert abs(float(O5.bce_dep_loss(b['x'], b['t'], o_m, z_m)) - b['loss']) < 1e-5 * abs(b['loss'])


def test_module_trees_match_the_reference_when_present(nets):
    """state_dict keys / shapes of the drop-in classes equal the reference's (build container only)."""
    from oracle import make_golden as MG
    if not MG.available():
        pytest.skip("reference tree not present")
    R = MG.load_ref("ref_networks5_ds_keys", "new_multi/networks5_ds.py")
    for name, args in (("G_1", {}), ("General_net", {}), ("R_dep", {}), ("_Discriminator", {"input_nc": 256})):
        ref = getattr(R, name)(**args).state_dict()
        got = getattr(nets, name)(**args).state_dict()
        assert list(ref.keys()) == list(got.keys()), name
        for k in ref:
            assert ref[k].shape == got[k].shape and ref[k].dtype == got[k].dtype, (name, k)
