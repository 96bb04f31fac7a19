"""ORACLE — test infrastructure only (imported by tests/ and oracle/make_golden.py; never by the product package).

Plain-PyTorch fp32 functional restatement of the U-Net task network of the reference's
``models/encoder_decoder.py``, driven by reference-layout ``state_dict``s:

* ``unet_encoder``   models/encoder_decoder.py:122-162 (+ _EncoderBlock :30-46, _InceptionBlock :47-82)
* ``unet_decoder``   models/encoder_decoder.py:164-208 (+ _DecoderUpBlock :84-101, _OutputBlock :103-117)
* ``unet_generator`` models/seg_network.py:441-559 (the two-headed U-Net of models/seg_model.py)
* ``SegCycleStepOracle``   models/seg_cycle.py:87-180 (Seg_basic, backward_G with the four task losses, ONE
  discriminator update per generator update)

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


def tie_prelu_prefixed(sd):
    """tie_prelu for a state_dict whose networks sit under prefixes ('scale0.'): ties within each prefix."""
    groups = {}
    for k, v in sd.items():
        if v.dim() == 1 and v.numel() == 1 and k.endswith('.weight'):
            groups.setdefault(k.split('.')[0], []).append(k)
    for keys in groups.values():
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


def unet_generator(sd, x, head, training=True, layers=4, weight=0.1):
    """models/seg_network.py:441-559 (_UNetGenerator.forward(input, syn_or_real)): the encoder of unet_encoder, the
    inception centre followed by the centre up-block (``center.<7-layers>``), then the ``*_syn`` (22 classes) or
    ``*_real`` (28 classes) decoder; every value other than 'syn' selects the real head, as the reference's else branch
    does.  Returns [center_in, output1].  The blocks are the same statements as models/encoder_decoder.py (the reference
    file repeats them, :155-285), so the helpers above are reused with this module's key names."""
    a = sd[ENC_SLOPE]
    conv1, conv2, conv3, center_in, cur = unet_encoder(sd, x, training, layers)
    center_out = _up_block(sd, 'center.%d.' % (7 - layers), cur, a, training)
    h = 'syn' if head == 'syn' else 'real'
    cat4 = torch.cat([center_out, conv3 * weight], 1)
    deconv4 = _up_block(sd, 'deconv4_%s.' % h, cat4, a, training)
    output4 = _out_block(sd, 'output4_%s.' % h, cat4)
    cat3 = torch.cat([deconv4, conv2 * weight * 0.5, _up2(output4)], 1)
    deconv3 = _up_block(sd, 'deconv3_%s.' % h, cat3, a, training)
    output3 = _out_block(sd, 'output3_%s.' % h, cat3)
    cat2 = torch.cat([deconv3, conv1 * weight * 0.1, _up2(output3)], 1)
    deconv2 = _up_block(sd, 'deconv2_%s.' % h, cat2, a, training)
    output2 = _out_block(sd, 'output2_%s.' % h, cat2)
    output1 = _out_block(sd, 'output1_%s.' % h, torch.cat([deconv2, _up2(output2)], 1))
    return [center_in, output1]


# ------------------------------------------------------------------------------------------------
# SegCycle training step (models/seg_cycle.py)
# ------------------------------------------------------------------------------------------------
from oracle.networks_oracle import CycleGANStepOracle, gan_loss, l1  # noqa: E402


class SegCycleStepOracle(CycleGANStepOracle):
    """Restated glue of SegCycle (models/seg_cycle.py:87-180): the CycleGAN step with four CrossEntropy(ignore 255)
    task losses added to loss_G (:129-136, :150-151) — (encoderA, decoderA) on real_A and (encoderB, decoderA) on
    fake_B against lab_A, (encoderB, decoderB) on real_B and (encoderA, decoderB) on fake_A against lab_B — Adam
    over the generators AND the task networks (:69-73) and a single discriminator update (:167-176)."""

    def __init__(self, sd_G_A, sd_G_B, sd_D_A, sd_D_B, sd_encA, sd_encB, sd_decA, sd_decB, lr=2e-4, beta1=0.5, **kw):
        super().__init__(sd_G_A, sd_G_B, sd_D_A, sd_D_B, lr=lr, beta1=beta1, d_iters=1, **kw)
        mk = lambda sd: tie_prelu({k: v.detach().clone().requires_grad_(v.is_floating_point() and 'running_' not in k)
                                   for k, v in sd.items()})
        self.encA, self.encB, self.decA, self.decB = mk(sd_encA), mk(sd_encB), mk(sd_decA), mk(sd_decB)
        seen, params = set(), []
        for d in (self.G_A, self.G_B, self.encA, self.encB, self.decA, self.decB):
            for p in d.values():
                if p.requires_grad and id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        self.opt_G = torch.optim.Adam(params, lr=lr, betas=(beta1, 0.999))

    def _seg(self, enc, dec, x, gt):
        out = unet_decoder(dec, unet_encoder(enc, x))
        return F.cross_entropy(out[-1], gt.squeeze(1), ignore_index=255)

    def step(self, real_A, real_B, lab_A, lab_B, train=True, apply_updates=True):
        fake_B = self._g(self.G_A, real_A)
        rec_A = self._g(self.G_B, fake_B)
        fake_A = self._g(self.G_B, real_B)
        rec_B = self._g(self.G_A, fake_A)
        self._set_requires_grad([self.D_A, self.D_B], False)
        self.opt_G.zero_grad()
        L = self.losses
        L['idt_A'] = l1(self._g(self.G_A, real_B), real_B) * self.lambda_B * self.lambda_idt
        L['idt_B'] = l1(self._g(self.G_B, real_A), real_A) * self.lambda_A * self.lambda_idt
        L['segAreal'] = self._seg(self.encA, self.decA, real_A, lab_A)
        L['segAfake'] = self._seg(self.encB, self.decA, fake_B, lab_A)
        L['segBreal'] = self._seg(self.encB, self.decB, real_B, lab_B)
        L['segBfake'] = self._seg(self.encA, self.decB, fake_A, lab_B)
        L['G_A'] = gan_loss(self._d(self.D_A, fake_B), True)
        L['G_B'] = gan_loss(self._d(self.D_B, fake_A), True)
        L['cycle_A'] = l1(rec_A, real_A) * self.lambda_A
        L['cycle_B'] = l1(rec_B, real_B) * self.lambda_B
        loss_G = (L['G_A'] + L['G_B'] + L['cycle_A'] + L['cycle_B'] + L['idt_A'] + L['idt_B'] + L['segAfake']
                  + L['segAreal'] + L['segBfake'] + L['segBreal'])
        L['G'] = loss_G
        if train:
            loss_G.backward()
            if apply_updates:
                self.opt_G.step()
        self._set_requires_grad([self.D_A, self.D_B], True)
        self.opt_D.zero_grad()
        L['D_A'] = self._d_loss(self.D_A, real_B, self.fake_B_pool.query(fake_B))
        L['D_B'] = self._d_loss(self.D_B, real_A, self.fake_A_pool.query(fake_A))
        if train:
            L['D_A'].backward()
            L['D_B'].backward()
            if apply_updates:
                self.opt_D.step()
        self.fake_A, self.fake_B, self.rec_A, self.rec_B = fake_A, fake_B, rec_A, rec_B
        return {k: float(v) for k, v in L.items()}
