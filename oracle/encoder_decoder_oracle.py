"""ORACLE — test infrastructure only (imported by tests/ and oracle/make_golden.py; never by the product package).

Plain-PyTorch fp32 functional restatement of the U-Net task network of the reference's
``models/encoder_decoder.py``, driven by reference-layout ``state_dict``s:

* ``unet_encoder``   models/encoder_decoder.py:122-162 (+ _EncoderBlock :30-46, _InceptionBlock :47-82)
* ``unet_decoder``   models/encoder_decoder.py:164-208 (+ _DecoderUpBlock :84-101, _OutputBlock :103-117)

The reference creates ONE ``nn.PReLU`` per network and hands it to every block (:134,:175): the slope is read from
the network's first PReLU key (``ENC_SLOPE`` / ``DEC_SLOPE``); ``tie_prelu`` makes a state_dict consistent with that.

Parity pin: the reference ships no tests or golden vectors.  This restatement is pinned against
the reference's OWN classes imported from /root/reference in the build container (tests/test_oracle_ed_pin.py) and
against outputs they generated (tests/golden/encoder_decoder.pt, written by oracle/make_golden.py).
"""
import torch
import torch.nn.functional as F

EPS = 1e-5
ENC_SLOPE = 'conv1.3.weight'
DEC_SLOPE = 'deconv_center.model.3.weight'


def tie_prelu(sd):
    """All PReLU keys (1-element tensors named *.weight) take the value of the network's first one."""
    keys = [k for k, v in sd.items() if v.dim() == 1 and v.numel() == 1 and k.endswith('.weight')
            and not k.endswith('num_batches_tracked')]
    for k in keys[1:]:
        sd[k] = sd[keys[0]]
    return sd


def _bn(x, sd, key, training):
    if key + '.weight' not in sd:           # InstanceNorm2d(affine=False)
        return F.instance_norm(x, eps=EPS)
    return F.batch_norm(x, sd.get(key + '.running_mean'), sd.get(key + '.running_var'), sd[key + '.weight'],
                        sd[key + '.bias'], training, 0.1, EPS)


def _rpad(x, p):
    return F.pad(x, (p, p, p, p), mode='reflect')


def _encoder_block(sd, p, x, a, training):
    h = F.prelu(_bn(F.conv2d(x, sd[p + 'model.0.weight'], sd.get(p + 'model.0.bias'), padding=1), sd, p + 'model.1',
                    training), a)
    return F.prelu(_bn(F.conv2d(h, sd[p + 'model.3.weight'], sd.get(p + 'model.3.bias'), padding=1), sd,
                       p + 'model.4', training), a)


def _inception(sd, p, x, a, width, training):
    branches = []
    for i in range(width):
        d = 2 * i + 1
        branches.append(F.conv2d(_rpad(x, d), sd['%slayer%d.1.weight' % (p, i)], sd.get('%slayer%d.1.bias' % (p, i)),
                                 dilation=d))
    h = F.prelu(_bn(torch.cat(branches, 1), sd, p + 'norm1', training), a)
    h = _bn(F.conv2d(_rpad(h, 1), sd[p + 'branch1x1.1.weight'], sd.get(p + 'branch1x1.1.bias')), sd, p + 'norm2',
            training)
    return F.prelu(h + x, a)


def unet_encoder(sd, x, training=True, layers=4):
    a = sd[ENC_SLOPE]
    h = F.conv2d(_rpad(x, 3), sd['conv1.1.weight'], sd.get('conv1.1.bias'))
    conv1 = F.avg_pool2d(F.prelu(_bn(h, sd, 'conv1.2', training), a), 2, 2)
    conv2 = F.avg_pool2d(_encoder_block(sd, 'conv2.', conv1, a, training), 2, 2)
    conv3 = F.avg_pool2d(_encoder_block(sd, 'conv3.', conv2, a, training), 2, 2)
    center_in = F.avg_pool2d(_encoder_block(sd, 'conv4.', conv3, a, training), 2, 2)
    cur = center_in
    for i in range(7 - layers):
        cur = _inception(sd, 'center.%d.' % i, cur, a, 7 - layers, training)
    return [conv1, conv2, conv3, center_in, cur]


def _up_block(sd, p, x, a, training):
    h = F.prelu(_bn(F.conv2d(_rpad(x, 1), sd[p + 'model.1.weight'], sd.get(p + 'model.1.bias')), sd, p + 'model.2',
                    training), a)
    h = F.conv_transpose2d(h, sd[p + 'model.4.weight'], sd[p + 'model.4.bias'], stride=2, padding=1, output_padding=1)
    return F.prelu(_bn(h, sd, p + 'model.5', training), a)


def _out_block(sd, p, x):
    w = sd[p + 'model.1.weight']
    return torch.tanh(F.conv2d(_rpad(x, w.shape[2] // 2), w, sd.get(p + 'model.1.bias')))


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode='nearest')


def unet_decoder(sd, feats, training=True, weight=0.1):
    conv1, conv2, conv3, center_in, center_out = feats
    a = sd[DEC_SLOPE]
    center = _up_block(sd, 'deconv_center.', center_out, a, training)
    cat4 = torch.cat([center, conv3 * weight], 1)
    deconv4 = _up_block(sd, 'deconv4.', cat4, a, training)
    output4 = _out_block(sd, 'output4.', cat4)
    cat3 = torch.cat([deconv4, conv2 * weight * 0.5, _up2(output4)], 1)
    deconv3 = _up_block(sd, 'deconv3.', cat3, a, training)
    output3 = _out_block(sd, 'output3.', cat3)
    cat2 = torch.cat([deconv3, conv1 * weight * 0.1, _up2(output3)], 1)
    deconv2 = _up_block(sd, 'deconv2.', cat2, a, training)
    output2 = _out_block(sd, 'output2.', cat2)
    output1 = _out_block(sd, 'output1.', torch.cat([deconv2, _up2(output2)], 1))
    return [center_in, output4, output3, output2, output1]
