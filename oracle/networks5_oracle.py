"""ORACLE — test infrastructure only (imported by tests/ and, for the cpu_baseline leg, bench.py; never by
the product package).

Plain-PyTorch fp32 functional restatement of the live networks of the reference's
``new_multi/networks5_ds.py`` and of the seg/depth training step of ``new_multi/model5.py``, driven by
reference-layout ``state_dict``s (keys WITHOUT the DataParallel ``module.`` prefix):

* ``g_1``            networks5_ds.py:26-66  (+ _DenseLayer/_DenseBlock :122-146, ResnetBlock :290-338)
* ``general_net``    networks5_ds.py:366-477 (+ _pspTrans :344-361); returns (head, [4 detached block outputs])
* ``r_dep``          networks5_ds.py:733-821 (+ G_side :623-651, seg_block :708-728, depth_block :655-706)
* ``discriminator``  networks5_ds.py:527-566 (shared PReLU at model.1 / model.10)
* ``bce_dep_loss`` / ``get_masks``   networks5_ds.py:947-956 / :973-982
* ``SegDepthStepOracle``   model5.py:415-696 (backward_G_2, backward_G_1, backward_R_D, backward_DISDEP)

Parity pin: the reference ships no tests or golden vectors.  The restatement is pinned
against the reference's OWN classes imported from /root/reference in the build container
(tests/test_oracle_pin.py) and against outputs they generated (tests/golden/networks5.pt, written by
oracle/make_golden.py).  The arithmetic itself is torch's (requirements.txt: torch>=0.4.0, unpinned).
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn(x, sd, key, training=True):
    return F.batch_norm(x, sd.get(key + '.running_mean'), sd.get(key + '.running_var'), sd[key + '.weight'],
                        sd[key + '.bias'], training, 0.1, EPS)


def _dense_block(sd, prefix, x, n_layers, training):
    """_DenseBlock: each layer is BN-ReLU-conv1x1-BN-ReLU-conv3x3 and concatenates its 32 new channels."""
    for i in range(1, n_layers + 1):
        p = '%sdenselayer%d.' % (prefix, i)
        h = F.relu(_bn(x, sd, p + 'norm1', training))
        h = F.conv2d(h, sd[p + 'conv1.weight'])
        h = F.relu(_bn(h, sd, p + 'norm2', training))
        h = F.conv2d(h, sd[p + 'conv2.weight'], padding=1)
        x = torch.cat([x, h], 1)
    return x


def _res_block(sd, p, x, training):
    """networks5_ds.ResnetBlock: x + BN(conv1x1_dil2(x)) + ReLU(BN(conv3x3(reflect_pad1(x))))."""
    a = _bn(F.conv2d(x, sd[p + 'conv0_block.1.weight'], sd.get(p + 'conv0_block.1.bias'), dilation=2), sd,
            p + 'conv0_block.2', training)
    b = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), sd[p + 'conv1_block.1.weight'],
                 sd.get(p + 'conv1_block.1.bias'))
    b = F.relu(_bn(b, sd, p + 'conv1_block.2', training))
    return x + a + b


def g_1(sd, x, training=True, n_blocks=3, n_layers=6):
    h = F.conv2d(x, sd['features.conv0.weight'], stride=2, padding=3)
    h = F.relu(_bn(h, sd, 'features.norm0', training))
    h = _dense_block(sd, 'features.denseblock1.', h, n_layers, training)
    h = F.conv2d(F.pad(h, (1, 1, 1, 1), mode='reflect'), sd['model.1.weight'], sd.get('model.1.bias'))
    h = F.relu(_bn(h, sd, 'model.2', training))
    for i in range(n_blocks):
        h = _res_block(sd, 'model.%d.' % (4 + i), h, training)
    return h


def _psp_trans(sd, p, x, training):
    h = F.relu(_bn(x, sd, p + 'trans.0', training))
    h = torch.cat([F.conv2d(h, sd[p + 'trans.2.weight']), F.conv2d(h, sd[p + 'trans.3.weight'], padding=1)], 1)
    return F.avg_pool2d(h, 2, 2)


def general_net(sd, x, kind, training=True, block_config=(6, 12, 32, 32)):
    if kind == 'R':
        x = F.conv2d(x, sd['features.conv0.weight'], stride=2, padding=3)
        x = F.relu(_bn(x, sd, 'features.norm0', training))
    feats = []
    for i, n_layers in enumerate(block_config):
        x = _dense_block(sd, 'features.denseblock%d.' % (i + 1), x, n_layers, training)
        feats.append(x.detach())
        if i < len(block_config) - 1:
            x = _psp_trans(sd, 'PSP.%d.' % i, x, training)
    h = F.relu(_bn(x, sd, 'psp.0', training))
    h = torch.cat([F.conv2d(h, sd['psp.2.weight']), F.conv2d(h, sd['psp.3.weight']),
                   F.conv2d(h, sd['psp.4.weight'], padding=1, dilation=2),
                   F.conv2d(h, sd['psp.5.weight'], padding=2, dilation=2)], 1)
    return _bn(h, sd, 'psp.6', training), feats


def _lrelu(x):
    return F.leaky_relu(x, 0.02)


def _g_side(sd, p, s, d, training):
    at = F.conv2d(d, sd[p + 'attention_bs.0.weight'], sd[p + 'attention_bs.0.bias'], stride=2, padding=1)
    at = F.adaptive_avg_pool2d(_lrelu(_bn(at, sd, p + 'attention_bs.1', training)), 1)
    sf = _bn(_lrelu(F.conv2d(s, sd[p + 'side_conv.0.weight'], sd[p + 'side_conv.0.bias'], padding=1)), sd,
             p + 'side_conv.2', training)
    sf = _bn(_lrelu(F.conv2d(sf, sd[p + 'side_conv.3.weight'], sd[p + 'side_conv.3.bias'])), sd, p + 'side_conv.5',
             training)
    out = d + torch.sigmoid(at) * sf
    out = _bn(_lrelu(F.conv2d(out, sd[p + 'conv.0.weight'], sd[p + 'conv.0.bias'], padding=1)), sd, p + 'conv.2',
              training)
    out = _bn(_lrelu(F.conv2d(out, sd[p + 'conv.3.weight'], sd[p + 'conv.3.bias'])), sd, p + 'conv.5', training)
    return F.interpolate(out, scale_factor=2, mode='bilinear', align_corners=True)


def _seg_block(sd, p, x, training):
    h = _bn(_lrelu(F.conv2d(x, sd[p + 'deconv.0.weight'], sd[p + 'deconv.0.bias'], padding=1)), sd, p + 'deconv.2',
            training)
    h = F.conv2d(h, sd[p + 'deconv.3.weight'], sd[p + 'deconv.3.bias'])
    h = F.interpolate(h, scale_factor=2, mode='bilinear', align_corners=True)
    return _lrelu(_bn(h, sd, p + 'bn', training))


def _depth_block(sd, p, x, training):
    dep_o, out_f = [], []
    for i in range(4):
        u = p + 'upconv.%d.' % i
        f = F.conv_transpose2d(x, sd[u + '0.weight'], sd[u + '0.bias'], stride=2, padding=1)
        f = _bn(_lrelu(f), sd, u + '2', training)
        f = _bn(_lrelu(F.conv2d(f, sd[u + '3.weight'], sd[u + '3.bias'], padding=1)), sd, u + '5', training)
        o = p + 'depth_out.%d.' % i
        dep_o.append(torch.tanh(F.conv2d(f, sd[o + '0.weight'], sd[o + '0.bias'], padding=1)))
        a = p + 'attention_bs.%d.' % i
        at = _bn(_lrelu(F.conv2d(x, sd[a + '0.weight'], sd[a + '0.bias'], stride=2, padding=1)), sd, a + '2', training)
        at = F.adaptive_avg_pool2d(_lrelu(F.conv2d(at, sd[a + '3.weight'], sd[a + '3.bias'], stride=2, padding=1)), 1)
        out_f.append(torch.sigmoid(at) * f + f)
    h = torch.cat(out_f, 1)
    h = _bn(_lrelu(F.conv2d(h, sd[p + 'conv.0.weight'], sd[p + 'conv.0.bias'], padding=1)), sd, p + 'conv.2', training)
    h = _bn(_lrelu(F.conv2d(h, sd[p + 'conv.3.weight'], sd[p + 'conv.3.bias'], padding=1)), sd, p + 'conv.5', training)
    dep1 = _bn(F.conv2d(h, sd[p + 'depconv.0.weight'], sd[p + 'depconv.0.bias'], padding=1), sd, p + 'depconv.1',
               training)
    return dep_o, dep1


def r_dep(sd, s_features, d_feature, training=True):
    out0 = _g_side(sd, 'AT.0.', s_features[3], d_feature, training)
    out1 = _g_side(sd, 'AT.1.', s_features[2], out0, training)
    out2 = _g_side(sd, 'AT.2.', s_features[1], out1, training)
    return (out0, out1, out2), _seg_block(sd, 'seg_d.', out2, training), _depth_block(sd, 'dep.', out2, training)


def discriminator(sd, x, training=True):
    """_Discriminator: conv-PReLU(shared)-[conv-BN-PReLU]x2-conv-BN-PReLU(shared)-conv."""
    shared = sd['model.1.weight']
    h = F.prelu(F.conv2d(x, sd['model.0.weight'], sd.get('model.0.bias'), stride=2, padding=1), shared)
    h = F.prelu(_bn(F.conv2d(h, sd['model.2.weight'], sd.get('model.2.bias'), stride=2, padding=1), sd, 'model.3',
                    training), sd['model.4.weight'])
    h = F.prelu(_bn(F.conv2d(h, sd['model.5.weight'], sd.get('model.5.bias'), stride=2, padding=1), sd, 'model.6',
                    training), sd['model.7.weight'])
    h = F.prelu(_bn(F.conv2d(h, sd['model.8.weight'], sd.get('model.8.bias'), stride=1, padding=1), sd, 'model.9',
                    training), shared)
    return F.conv2d(h, sd['model.11.weight'], sd['model.11.bias'], stride=1, padding=1)


def get_masks(target):
    o_m = (target == 1).to(target.dtype)
    z_m = (target == -1).to(target.dtype)
    return o_m, z_m


def bce_dep_loss(x, target, o_m, z_m):
    if torch.is_autocast_enabled():     # F.binary_cross_entropy refuses to run under autocast (envelope runs only)
        with torch.autocast('cuda', enabled=False):
            return bce_dep_loss(x.float(), target.float(), o_m.float(), z_m.float())
    return (F.binary_cross_entropy((x + 1) / 2 * o_m, (target + 1) / 2 * o_m)
            + F.binary_cross_entropy((x + 1) / 2 * z_m, (target + 1) / 2 * z_m) + 50 * F.l1_loss(x.expand_as(target), target))


def gan_mse(pred, is_real):
    return F.mse_loss(pred, torch.full_like(pred, 1.0 if is_real else 0.0))


def strip_module_prefix(sd):
    return {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}


def leaf_params(sd):
    """Copies of a state_dict whose parameters (not the BatchNorm buffers) require grad."""
    return {k: v.detach().clone().requires_grad_(v.is_floating_point() and 'running_' not in k) for k, v in sd.items()}


# ------------------------------------------------------------------------------------------------
# seg/depth training step (new_multi/model5.py:415-696)
# ------------------------------------------------------------------------------------------------
class SegDepthStepOracle:
    """Restated glue of Seg_Depth.optimize_parameters (:640-696): backward_G_2 (:585-638) -> step G_2;
    backward_G_1 (:564-583) -> step G_1; backward_R_D (:479-559) -> two R_D steps; backward_DISDEP (:415-474)
    -> three feature-discriminator steps.  Adam learning rates lr/5 (G_1), lr/3 (G_2), lr/2 (R_D), lr/4 (FD*)
    with betas (beta1, 0.999) (:250-275).  The reference's quirks are kept: L1 between a [B,1,H,W] prediction
    and a [B,H,W] label broadcasts to [B,B,H,W] (:532,573,608), FD3 is left trainable for backward_R_D's GAN
    terms only through its inputs (requires_grad toggles :586-590,653-693)."""

    def __init__(self, sd_G1, sd_G2, sd_RD, sd_FD1, sd_FD2, sd_FD3, lr=2e-4, beta1=0.5):
        self.G1, self.G2, self.RD = leaf_params(sd_G1), leaf_params(sd_G2), leaf_params(sd_RD)
        self.FD = [leaf_params(sd_FD1), leaf_params(sd_FD2), leaf_params(sd_FD3)]
        ps = lambda d: list({id(p): p for p in d.values() if p.requires_grad}.values())
        self._ps = ps
        adam = lambda d, f: torch.optim.Adam(ps(d), lr=lr / f, betas=(beta1, 0.999))
        self.opt_G1, self.opt_G2, self.opt_RD = adam(self.G1, 5), adam(self.G2, 3), adam(self.RD, 2)
        self.opt_FD = [adam(d, 4) for d in self.FD]
        for d in self.FD:    # model.10.weight is model.1.weight (one shared nn.PReLU)
            d['model.10.weight'] = d['model.1.weight']
        self._all = {'G1': ps(self.G1), 'G2': ps(self.G2), 'RD': ps(self.RD), 'FD1': ps(self.FD[0]),
                     'FD2': ps(self.FD[1]), 'FD3': ps(self.FD[2])}
        self.losses = {}

    def _req(self, names, flag):
        for n in names:
            for p in self._all[n]:
                p.requires_grad_(flag)

    @staticmethod
    def _ce(seg, lab):
        return F.cross_entropy(seg, lab, ignore_index=255)

    @staticmethod
    def _sky(seg_l):
        m = seg_l.clone()
        m[m != 17] = 1
        m[seg_l == 17] = 0
        return m

    def step(self, syn_img, real_img, syn_seg_l, real_seg_l, syn_dep_l, syn_dep_ls, apply_updates=True):
        L = self.losses
        # ---------------- backward_G_2 (:585-638)
        self._req(['G2'], True)
        self.opt_G2.zero_grad()
        self._req(['RD', 'G1', 'FD1', 'FD2'], False)
        ss = g_1(self.G1, syn_img)
        head, feats = general_net(self.G2, ss.detach(), 'S')
        _, seg, (dep_4, dep_o) = r_dep(self.RD, feats, head)
        sky = self._sky(syn_seg_l)
        dep_loss = F.l1_loss(dep_o, sky.float() * syn_dep_l)
        loss_syn = dep_loss + self._ce(seg, syn_seg_l)
        syn_head, syn_feats = head.detach(), feats
        rhead, rfeats = general_net(self.G2, real_img, 'R')
        _, rseg, _ = r_dep(self.RD, rfeats, rhead)
        real_head, real_feats = rhead.detach(), rfeats
        L['G2'] = loss_syn + 2 * self._ce(rseg, real_seg_l)
        L['G2'].backward()
        if apply_updates:
            self.opt_G2.step()
        # ---------------- backward_G_1 (:564-583)
        self._req(['G1'], True)
        self._req(['G2'], False)
        self.opt_G1.zero_grad()
        ss = g_1(self.G1, syn_img)
        h1, f1 = general_net(self.G2, ss, 'S')
        _, s_seg, (_, s_dep_o) = r_dep(self.RD, f1, h1)
        L['G1'] = self._ce(s_seg, syn_seg_l) + F.l1_loss(s_dep_o, syn_dep_l)
        L['G1'].backward()
        if apply_updates:
            self.opt_G1.step()
        # ---------------- backward_R_D (:479-559)
        self._req(['G1', 'G2'], False)
        self._req(['RD'], True)
        self.opt_RD.zero_grad()
        feats_r, seg_r, (_, dep_o_r) = r_dep(self.RD, real_feats, real_head)
        preds = [discriminator(self.FD[i], feats_r[i]) for i in range(3)]
        L['RD_real'] = self._ce(seg_r, real_seg_l) + sum(0.2 * gan_mse(p, False) for p in preds)
        L['RD_real'].backward()
        if apply_updates:
            self.opt_RD.step()
        real_feats_out = [f.detach() for f in feats_r]
        self.opt_RD.zero_grad()
        feats_s, seg_s, (dep_4, dep_o) = r_dep(self.RD, syn_feats, syn_head)
        sky = self._sky(syn_seg_l)
        sky4 = torch.cat([sky.unsqueeze(1)] * 4, 1).float() * syn_dep_ls
        oms, zms = get_masks(sky4)
        dep_loss = F.l1_loss(dep_o, sky.float() * syn_dep_l)
        for s_dep in dep_4:
            dep_loss = dep_loss + bce_dep_loss(sky.unsqueeze(1).float() * s_dep, sky4, oms, zms)
        L['RD_syn'] = dep_loss + self._ce(seg_s, syn_seg_l)
        L['dep_ref'] = dep_loss.detach()
        L['RD_syn'].backward()
        if apply_updates:
            self.opt_RD.step()
        syn_feats_out = [f.detach() for f in feats_s]
        # ---------------- backward_DISDEP (:415-474)
        self._req(['G1', 'G2', 'RD'], False)
        self._req(['FD1', 'FD2', 'FD3'], True)
        for i in range(3):
            self.opt_FD[i].zero_grad()
            d_real = discriminator(self.FD[i], real_feats_out[i])
            d_fake = discriminator(self.FD[i], syn_feats_out[i])
            L['FD%d' % (i + 1)] = gan_mse(d_real, True) + gan_mse(d_fake, False)
            L['FD%d' % (i + 1)].backward()
            if apply_updates:
                self.opt_FD[i].step()
        self._req(['FD1', 'FD2', 'FD3'], False)
        self.syn_dep_ref, self.real_dep_ref = dep_o.squeeze(1).detach(), dep_o_r.squeeze(1).detach()
        self.real_feats, self.syn_feats = real_feats_out, syn_feats_out
        return {k: float(v.detach()) for k, v in L.items()}


# ------------------------------------------------------------------------------------------------
# deterministic weights keyed by parameter NAME (shared by make_golden.py and the tests, so that fixtures
# need not carry the 23 M / 53 M parameter state_dicts)
# ------------------------------------------------------------------------------------------------
def synth_state_dict(template, seed=0):
    """template: {key: tensor} (only shapes / dtypes are used). Conv / PReLU weights ~ N(0, s) with s chosen so
    activations stay O(1); BatchNorm weight ~ 1 + 0.1 N, bias ~ 0.1 N, running_mean ~ 0.1 N, running_var ~ 1 + 0.1 U."""
    import zlib
    out = {}
    for k, v in template.items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) % (2 ** 31))
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros_like(v)
        elif k.endswith('running_mean'):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('running_var'):
            out[k] = 1.0 + 0.1 * torch.rand(v.shape, generator=g)
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            out[k] = torch.randn(v.shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif v.dim() == 1 and v.numel() == 1:      # PReLU slope
            out[k] = torch.full(v.shape, 0.25) + 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith('.weight'):                # BatchNorm weight
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:                                      # biases
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
    return out
