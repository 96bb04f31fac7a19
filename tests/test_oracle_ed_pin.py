"""Pins oracle/encoder_decoder_oracle.py against the fixture the REFERENCE's own
models/encoder_decoder.py classes produced (oracle/make_golden.py -> tests/golden/encoder_decoder.pt; weights
regenerated here from names + seed), checks that the drop-in modules expose the reference's state_dict layout, and —
in the build container, where /root/reference exists — compares with the reference classes directly.  CPU only.
Tolerance: the same fp32 torch ops in a different composition -> 2e-5 relative L2."""
import importlib.util
import os

import pytest
import torch

from helpers import rel_l2
from oracle import encoder_decoder_oracle as OE
from oracle import networks5_oracle as O5

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference/models/encoder_decoder.py"
TOL = 2e-5


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLD, "encoder_decoder.pt"), weights_only=False)


def _nets(ngf=8, nc=5):
    from cycle_depth_estimation_b200 import encoder_decoder as E
    return E._UNetEncoder(input_nc=3, ngf=ngf), E._UNetDecoder(output_nc=nc, ngf=ngf)


def test_state_dict_layout_and_shared_prelu(fx):
    from cycle_depth_estimation_b200 import encoder_decoder as E
    enc, dec = _nets()
    assert list(enc.state_dict().keys()) == fx['enc_keys']
    assert list(dec.state_dict().keys()) == fx['dec_keys']
    # one nn.PReLU per network, visible under every block's key
    slopes = [k for k, v in enc.state_dict().items() if v.shape == (1,) and k.endswith('.weight')]
    assert slopes[0] == OE.ENC_SLOPE and len(slopes) == 1 + 2 * 3 + 3
    assert len({v.data_ptr() for k, v in enc.state_dict().items() if k in slopes}) == 1
    assert len(list(enc.parameters())) == len({id(p) for p in enc.parameters()})
    assert next(k for k, v in dec.state_dict().items() if v.shape == (1,) and k.endswith('.weight')) == OE.DEC_SLOPE
    with pytest.raises(RuntimeError):
        enc.conv2(torch.zeros(1, 8, 4, 4))       # parameter containers only: the tape runs the blocks
    with pytest.raises(NotImplementedError):
        E.get_norm_layer('group')
    with pytest.raises(NotImplementedError):
        E.get_nonlinearity_layer('GELU')
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 3, 96, 96))           # CPU tensors are refused (no CPU path)


def test_oracle_matches_reference_fixture(fx):
    enc, dec = _nets()
    sd_e = OE.tie_prelu(O5.leaf_params(OE.tie_prelu(O5.synth_state_dict(enc.state_dict(), 11))))
    sd_d = OE.tie_prelu(O5.leaf_params(OE.tie_prelu(O5.synth_state_dict(dec.state_dict(), 12))))
    x = fx['x'].clone().requires_grad_(True)
    feats = OE.unet_encoder(sd_e, x)
    outs = OE.unet_decoder(sd_d, feats)
    for a, b in zip(feats, fx['feats']):
        assert a.shape == b.shape and rel_l2(a, b) < TOL
    assert outs[0] is feats[3]
    for a, b in zip(outs[1:], fx['outs']):
        assert a.shape == b.shape and rel_l2(a, b) < TOL
    (outs[-1] * fx['gout']).sum().backward()
    assert rel_l2(x.grad, fx['gx']) < 1e-4
    assert rel_l2(sd_e[OE.ENC_SLOPE].grad, fx['g_enc_slope']) < 1e-4
    assert rel_l2(sd_d[OE.DEC_SLOPE].grad, fx['g_dec_slope']) < 1e-4
    assert rel_l2(sd_d['output1.model.1.weight'].grad, fx['g_dec_out1']) < 1e-4


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree exists only in the build container")
def test_oracle_matches_reference_classes_instance_norm():
    """norm='instance' (biases live, no affine) and eval-mode BatchNorm against the reference modules themselves."""
    from cycle_depth_estimation_b200 import encoder_decoder as E
    spec = importlib.util.spec_from_file_location("ref_encoder_decoder", REF)
    ED = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ED)
    g = torch.Generator().manual_seed(5)
    x = torch.rand((2, 3, 96, 128), generator=g) * 2 - 1
    for norm, training in (('instance', True), ('batch', False)):
        enc, dec = ED._UNetEncoder(3, ngf=8, norm=norm), ED._UNetDecoder(4, ngf=8, norm=norm)
        sd_e = OE.tie_prelu(O5.synth_state_dict(enc.state_dict(), 21))
        sd_d = OE.tie_prelu(O5.synth_state_dict(dec.state_dict(), 22))
        enc.load_state_dict(sd_e, strict=True)
        dec.load_state_dict(sd_d, strict=True)
        enc.train(training)
        dec.train(training)
        with torch.no_grad():
            want = dec(enc(x))
            got = OE.unet_decoder(sd_d, OE.unet_encoder(sd_e, x, training), training)
        for a, b in zip(got[1:], want[1:]):
            assert rel_l2(a, b) < TOL, (norm, rel_l2(a, b))
        mine_e, mine_d = E._UNetEncoder(3, ngf=8, norm=norm), E._UNetDecoder(4, ngf=8, norm=norm)
        assert list(mine_e.state_dict().keys()) == list(enc.state_dict().keys())
        assert ([tuple(v.shape) for v in mine_d.state_dict().values()]
                == [tuple(v.shape) for v in dec.state_dict().values()])
        mine_d.load_state_dict(dec.state_dict(), strict=True)
