"""Drop-in replacement for the live part of the reference's ``new_multi/networks5_ds.py``: the networks the
seg/depth step of ``new_multi/model5.py`` builds (:199-285) — ``G_1``, ``General_net``, ``R_dep``,
``_Discriminator`` — with ``GANLoss``, ``BCEDepLoss``, ``get_masks``, ``init_weights`` and ``init_net``.

Same constructor signatures, module tree and ``state_dict`` keys as the reference, so its checkpoints
(``new_multi/checkpoints/iter_4000_net_G_1.pth``) load strictly.  The leaf modules are stock ``torch.nn``
classes used as parameter holders; every ``forward`` runs on the graph engine (``graph.py``):

* dense blocks (:122-146) live in ONE NHWC buffer per block; each layer's 3x3 convolution writes its 32 new
  channels (and their batch statistics, from the convolution epilogue) at the next channel offset, the
  BN-ReLU-1x1 of a layer reads a channel prefix through a strided TMA map; ``torch.cat`` never runs and the
  per-channel batch sums are computed once instead of once per consuming layer;
* the gradient of a block buffer is one fp32 accumulator that every consumer's BatchNorm backward adds into;
* conv -> LeakyReLU -> BatchNorm blocks of ``R_dep`` apply the activation in the convolution epilogue;
* channel attention (GAP + sigmoid + mul + add), 2x2 average pooling, bilinear x2 (align_corners=True) and
  PReLU are single fused kernels with hand-written backward.

The classes of the reference file that ``model5`` never instantiates (``Discriminator``, ``SEG``, ``DEP``,
``_MultiscaleDiscriminator``, ``_FeatureDiscriminator``, ``Discriminator2_seg``, ``DenseNet``) are not part
of the hot path and are not provided.
"""
import functools
from collections import OrderedDict

import torch
import torch.nn as nn
from torch.nn import init

from . import engine, graph, losses
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_TANH


# ------------------------------------------------------------------------------------------------
# helpers (new_multi/networks5_ds.py:229-262, 478-500)
# ------------------------------------------------------------------------------------------------
def init_weights(net, init_type='normal', gain=0.02):
    """new_multi/networks5_ds.py:229-250."""
    fillers = {
        'normal': lambda w: init.normal_(w, 0.0, gain),
        'xavier': lambda w: init.xavier_normal_(w, gain=gain),
        'kaiming': lambda w: init.kaiming_normal_(w, a=0, mode='fan_in'),
        'orthogonal': lambda w: init.orthogonal_(w, gain=gain),
    }

    def visit(m):
        name = type(m).__name__
        if hasattr(m, 'weight') and ('Conv' in name or 'Linear' in name):
            if init_type not in fillers:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            fillers[init_type](m.weight.data)
            if getattr(m, 'bias', None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif 'BatchNorm2d' in name:
            init.normal_(m.weight.data, 1.0, gain)
            init.constant_(m.bias.data, 0.0)

    print('initialize network with %s' % init_type)
    net.apply(visit)
    engine.invalidate_packed_weights()


class _Replica(nn.Module):
    """What ``init_net`` returns in place of ``nn.DataParallel(net)`` (:258-259): a transparent wrapper that
    keeps the ``module.`` prefix of the state_dict keys (the shipped checkpoints carry it) but runs the
    network in this process — data parallelism is one process per GPU here."""

    def __init__(self, module):
        super(_Replica, self).__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def init_net(net, init_type='normal', init_gain=0.02):
    """new_multi/networks5_ds.py:252-262: .cuda(), DataParallel wrapper, init_weights."""
    net = _Replica(net.cuda())
    init_weights(net, init_type, gain=init_gain)
    return net


def get_norm_layer(norm_type='batch'):
    """new_multi/networks5_ds.py:478-487."""
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False)
    if norm_type == 'none':
        return None
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def get_nonlinearity_layer(activation_type='PReLU'):
    """new_multi/networks5_ds.py:490-501."""
    if activation_type == 'ReLU':
        return nn.ReLU(True)
    if activation_type == 'SELU':
        return nn.SELU(True)
    if activation_type == 'LeakyReLU':
        return nn.LeakyReLU(0.1, True)
    if activation_type == 'PReLU':
        return nn.PReLU()
    raise NotImplementedError('activation layer [%s] is not found' % activation_type)


def _bias_follows_instance_norm(norm_layer):
    target = norm_layer.func if type(norm_layer) == functools.partial else norm_layer
    return target == nn.InstanceNorm2d


def _parts_only(name):
    def forward(self, *a, **k):
        raise RuntimeError("%s runs as part of a fused cdb200 network" % name)
    return forward


# ------------------------------------------------------------------------------------------------
# DenseNet pieces (new_multi/networks5_ds.py:122-146)
# ------------------------------------------------------------------------------------------------
class _DenseLayer(nn.Sequential):
    def __init__(self, num_input_features, growth_rate, bn_size, drop_rate):
        super(_DenseLayer, self).__init__()
        self.add_module('norm1', nn.BatchNorm2d(num_input_features))
        self.add_module('relu1', nn.ReLU(inplace=True))
        self.add_module('conv1', nn.Conv2d(num_input_features, bn_size * growth_rate, kernel_size=1, stride=1,
                                           bias=False))
        self.add_module('norm2', nn.BatchNorm2d(bn_size * growth_rate))
        self.add_module('relu2', nn.ReLU(inplace=True))
        self.add_module('conv2', nn.Conv2d(bn_size * growth_rate, growth_rate, kernel_size=3, stride=1, padding=1,
                                           bias=False))
        self.drop_rate = drop_rate
        if drop_rate > 0:
            raise NotImplementedError("dense-layer dropout (the reference always builds drop_rate=0)")

    forward = _parts_only('_DenseLayer')


class _DenseBlock(nn.Sequential):
    def __init__(self, num_layers, num_input_features, bn_size, growth_rate, drop_rate):
        super(_DenseBlock, self).__init__()
        for i in range(num_layers):
            self.add_module('denselayer%d' % (i + 1),
                            _DenseLayer(num_input_features + i * growth_rate, growth_rate, bn_size, drop_rate))
        self.num_input_features = num_input_features
        self.out_features = num_input_features + num_layers * growth_rate

    forward = _parts_only('_DenseBlock')


def _dense_block(tape, block, buf):
    """Runs a dense block in place on its concatenation buffer ``buf`` whose first channels (the block
    input) and their statistics are already filled in (new_multi/networks5_ds.py:132-146)."""
    k = block.num_input_features
    for layer in block.children():
        a = tape.norm_act(buf.slice(0, k), layer.norm1, ACT_RELU)
        b = tape.stage(a, layer.conv1, layer.norm2, ACT_RELU, halo=1, halo_kind='zero')
        g = layer.conv2.out_channels
        tape.stage(b, layer.conv2, None, ACT_NONE, out=buf.slice(k, k + g))
        k += g
    return buf


class _pspTrans(nn.Module):
    """new_multi/networks5_ds.py:344-361: BN-ReLU-[1x1 || 3x3]-cat-AvgPool2."""

    def __init__(self, num_input_features):
        super(_pspTrans, self).__init__()
        q = int(num_input_features / 4)
        self.trans = nn.ModuleList([
            nn.BatchNorm2d(num_input_features), nn.ReLU(inplace=False),
            nn.Conv2d(num_input_features, q, kernel_size=1, stride=1, bias=False),
            nn.Conv2d(num_input_features, q, kernel_size=3, stride=1, padding=1, bias=False),
            nn.AvgPool2d(kernel_size=2, stride=2)])

    forward = _parts_only('_pspTrans')


def _psp_trans(tape, m, buf, out):
    n, h, w, _ = buf.t.shape
    t = tape.norm_act(buf, m.trans[0], ACT_RELU, halo=1, halo_kind='zero')
    q = m.trans[2].out_channels
    cat = tape.concat_buffer(n, h, w, 2 * q)
    tape.stage(t, m.trans[2], None, ACT_NONE, out=cat.slice(0, q))
    tape.stage(t, m.trans[3], None, ACT_NONE, out=cat.slice(q, 2 * q))
    return tape.avgpool2(cat, out=out)


class ResnetBlock(nn.Module):
    """new_multi/networks5_ds.py:290-338: x + BN(conv1x1(x)) + ReLU(BN(conv3x3(reflect_pad(x))))."""

    def __init__(self, in_dim, padding_type, norm_layer, use_dropout, use_bias):
        super(ResnetBlock, self).__init__()
        if padding_type not in ('reflect', 'replicate', 'zero'):
            raise NotImplementedError('padding [%s] is not implemented' % padding_type)
        if padding_type == 'replicate':
            raise NotImplementedError('cdb200: replicate padding')
        block0 = [nn.ReflectionPad2d(0)] if padding_type == 'reflect' else []
        block0 += [nn.Conv2d(in_dim, in_dim, kernel_size=1, dilation=2, padding=0, bias=use_bias), norm_layer(in_dim)]
        block1 = [nn.ReflectionPad2d(1), nn.Conv2d(in_dim, in_dim, kernel_size=3, bias=use_bias), norm_layer(in_dim),
                  nn.ReLU(False)]
        if use_dropout:
            raise NotImplementedError("cdb200: dropout inside networks5_ds.ResnetBlock")
        self.conv0_block = nn.Sequential(*block0)
        self.conv1_block = nn.Sequential(*block1)

    forward = _parts_only('ResnetBlock')


def _resnet_block(tape, m, x, halo, halo_kind):
    c0 = [l for l in m.conv0_block if isinstance(l, nn.Conv2d)][0]
    n0 = [l for l in m.conv0_block if isinstance(l, (nn.BatchNorm2d, nn.InstanceNorm2d))][0]
    c1 = [l for l in m.conv1_block if isinstance(l, nn.Conv2d)][0]
    n1 = [l for l in m.conv1_block if isinstance(l, (nn.BatchNorm2d, nn.InstanceNorm2d))][0]
    a = tape.stage(x, c0, n0, ACT_NONE, res=x)
    return tape.stage(x, c1, n1, ACT_RELU, reflect=1, res=a, halo=halo, halo_kind=halo_kind)


# ------------------------------------------------------------------------------------------------
# G_1 (new_multi/networks5_ds.py:26-66)
# ------------------------------------------------------------------------------------------------
class G_1(nn.Module):
    def __init__(self, growth_rate=32, block_config=6, num_init_features=64, bn_size=4, drop_rate=0, ngf=64,
                 norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=3, padding_type='reflect'):
        assert (n_blocks > 0)
        super(G_1, self).__init__()
        self.input_nc = 3
        use_bias = _bias_follows_instance_norm(norm_layer)
        self.features = nn.Sequential(OrderedDict([
            ('conv0', nn.Conv2d(3, num_init_features, kernel_size=7, stride=2, padding=3, bias=False)),
            ('norm0', nn.BatchNorm2d(num_init_features)),
            ('relu0', nn.ReLU(inplace=False)),
        ]))
        self.features.add_module('denseblock1', _DenseBlock(num_layers=block_config,
                                                            num_input_features=num_init_features, bn_size=bn_size,
                                                            growth_rate=growth_rate, drop_rate=drop_rate))
        num_features = num_init_features + block_config * growth_rate
        model = [nn.ReflectionPad2d(1), nn.Conv2d(num_features, ngf, kernel_size=3, padding=0, bias=use_bias),
                 norm_layer(ngf), nn.ReLU(False)]
        model += [ResnetBlock(ngf, padding_type=padding_type, norm_layer=norm_layer, use_dropout=use_dropout,
                              use_bias=use_bias) for _ in range(n_blocks)]
        self.model = nn.Sequential(*model)

    def _body(self, tape, x):
        f = self.features
        img = tape.input_nchw(x, first_conv=f.conv0, want_grad=tape.input_wants[0])
        n = x.shape[0]
        ho, wo = tape._out_hw(f.conv0, False, x.shape[2], x.shape[3], 0)
        block = f.denseblock1
        buf = tape.concat_buffer(n, ho, wo, block.out_features, f32grad=True, stats=True)
        tape.stage(img, f.conv0, f.norm0, ACT_RELU, out=buf.slice(0, block.num_input_features))
        _dense_block(tape, block, buf)
        mods = list(self.model.children())
        padded = tape.norm_act(buf, None, ACT_NONE, halo=1, halo_kind='reflect')     # nn.ReflectionPad2d(1)
        blocks = [m for m in mods if isinstance(m, ResnetBlock)]
        cur = tape.stage(padded, mods[1], mods[2], ACT_RELU, reflect=1, halo=1, halo_kind='reflect')
        for i, m in enumerate(blocks):
            last = i == len(blocks) - 1
            cur = _resnet_block(tape, m, cur, 0 if last else 1, None if last else 'reflect')
        out, slot = tape.output_nchw(cur)
        return [out], [slot], [img]

    def forward(self, input):
        return graph.run(self, self._body, [input])[0]


# ------------------------------------------------------------------------------------------------
# General_net (new_multi/networks5_ds.py:366-477)
# ------------------------------------------------------------------------------------------------
class General_net(nn.Module):
    def __init__(self, mid_nc=1024, num_init_features=64, growth_rate=32, block_config=(6, 12, 32, 32), bn_size=4,
                 drop_rate=0):
        super(General_net, self).__init__()
        self.features = nn.Sequential(OrderedDict([
            ('conv0', nn.Conv2d(3, num_init_features, kernel_size=7, stride=2, padding=3, bias=False)),
            ('norm0', nn.BatchNorm2d(num_init_features)),
            ('relu0', nn.ReLU(inplace=True)),
        ]))
        num_features = num_init_features
        self.PSP = nn.ModuleList()
        for i, num_layers in enumerate(block_config):
            self.features.add_module('denseblock%d' % (i + 1),
                                     _DenseBlock(num_layers=num_layers, num_input_features=num_features,
                                                 bn_size=bn_size, growth_rate=growth_rate, drop_rate=drop_rate))
            num_features = num_features + num_layers * growth_rate
            if i != len(block_config) - 1:
                self.PSP.append(_pspTrans(num_features))
                num_features = num_features // 2
        q = int(mid_nc / 4)
        self.psp = nn.ModuleList([
            nn.BatchNorm2d(num_features), nn.ReLU(inplace=True),
            nn.Conv2d(num_features, q, kernel_size=1, stride=1, bias=False),
            nn.Conv2d(num_features, q, kernel_size=1, stride=1, dilation=1, bias=False),
            nn.Conv2d(num_features, q, kernel_size=2, stride=1, padding=1, dilation=2, bias=False),
            nn.Conv2d(num_features, q, kernel_size=3, stride=1, padding=2, dilation=2, bias=False),
            nn.BatchNorm2d(mid_nc)])

    def _body(self, kind, tape, x):
        f = self.features
        blocks = [m for m in f.children() if isinstance(m, _DenseBlock)]
        n = x.shape[0]
        if kind == 'R':
            img = tape.input_nchw(x, first_conv=f.conv0, want_grad=tape.input_wants[0])
            h, w = tape._out_hw(f.conv0, False, x.shape[2], x.shape[3], 0)
            buf = tape.concat_buffer(n, h, w, blocks[0].out_features, f32grad=True, stats=True)
            tape.stage(img, f.conv0, f.norm0, ACT_RELU, out=buf.slice(0, blocks[0].num_input_features))
        else:
            h, w = x.shape[2], x.shape[3]
            if x.shape[1] != blocks[0].num_input_features:
                raise RuntimeError("General_net('S') expects %d input channels" % blocks[0].num_input_features)
            buf = tape.concat_buffer(n, h, w, blocks[0].out_features, f32grad=True, stats=True)
            img = tape.input_nchw(x, want_grad=tape.input_wants[0], out=buf.slice(0, x.shape[1]))
        outs, slots = [], []
        feats = []
        for i, block in enumerate(blocks):
            _dense_block(tape, block, buf)
            feats.append(buf)
            if i < len(blocks) - 1:
                h, w = h // 2, w // 2
                nxt = tape.concat_buffer(n, h, w, blocks[i + 1].out_features, f32grad=True, stats=True)
                _psp_trans(tape, self.PSP[i], buf, nxt.slice(0, blocks[i + 1].num_input_features))
                buf = nxt
        u = tape.norm_act(buf, self.psp[0], ACT_RELU)
        q = self.psp[2].out_channels
        cat = tape.concat_buffer(n, h, w, 4 * q)
        for j in range(4):
            tape.stage(u, self.psp[2 + j], None, ACT_NONE, out=cat.slice(j * q, (j + 1) * q))
        head = tape.norm_act(cat, self.psp[6], ACT_NONE)
        o, s = tape.output_nchw(head)
        outs.append(o)
        slots.append(s)
        for fb in feats:   # features.append(input.detach()) (:427,:459)
            o, s = tape.output_nchw(fb, differentiable=False)
            outs.append(o)
            slots.append(s)
        return outs, slots, [img]

    def forward(self, input, type):
        if type not in ('R', 'S'):
            return None     # the reference falls off the end of forward for any other value
        outs = graph.run(self, functools.partial(self._body, type), [input])
        return outs[0], [t.detach() for t in outs[1:]]


# ------------------------------------------------------------------------------------------------
# R_dep and its blocks (new_multi/networks5_ds.py:68-87, 623-821)
# ------------------------------------------------------------------------------------------------
class DeconvBlock(torch.nn.Module):
    """new_multi/networks5_ds.py:68-87 (R_dep.up0: present in the state_dict, never called by R_dep.forward)."""

    def __init__(self, input_size, output_size, kernel_size=4, stride=2, padding=1, batch_norm=False, dropout=False):
        super(DeconvBlock, self).__init__()
        self.deconv = torch.nn.ConvTranspose2d(input_size, output_size, kernel_size, stride, padding)
        self.bn = torch.nn.BatchNorm2d(output_size)
        self.relu = torch.nn.LeakyReLU(0.02)
        self.batch_norm = batch_norm
        self.dropout = dropout

    forward = _parts_only('DeconvBlock')


def _conv_act_bn_chain(tape, seq, x, halo=0, halo_kind=None):
    """Runs an nn.Sequential made of [Conv|ConvT, LeakyReLU, BatchNorm] triples (activation BEFORE the norm),
    [Conv, BatchNorm, LeakyReLU] triples, bare convs, a trailing nn.UpsamplingBilinear2d / AdaptiveAvgPool2d is
    left to the caller. Returns the last value."""
    mods = [m for m in seq.children()]
    i, cur = 0, x
    convs = (nn.Conv2d, nn.ConvTranspose2d)
    while i < len(mods):
        m = mods[i]
        if not isinstance(m, convs):
            break
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        nxt2 = mods[i + 2] if i + 2 < len(mods) else None
        # does a stride-1 3x3 conv follow? then materialise its zero padding
        j = i + 1
        while j < len(mods) and not isinstance(mods[j], convs):
            j += 1
        follow = mods[j] if j < len(mods) else None
        h, hk = 0, None
        if isinstance(follow, nn.Conv2d) and follow.stride[0] == 1 and follow.padding[0] > 0 and follow.dilation[0] == 1:
            h, hk = follow.padding[0], 'zero'
        elif follow is None:
            h, hk = halo, halo_kind
        if isinstance(nxt, nn.LeakyReLU) and isinstance(nxt2, nn.BatchNorm2d):
            cur = tape.stage(cur, m, nxt2, ACT_LEAKY, float(nxt.negative_slope), act_first=True, halo=h, halo_kind=hk)
            i += 3
        elif isinstance(nxt, nn.BatchNorm2d) and isinstance(nxt2, nn.LeakyReLU):
            cur = tape.stage(cur, m, nxt, ACT_LEAKY, float(nxt2.negative_slope), halo=h, halo_kind=hk)
            i += 3
        elif isinstance(nxt, nn.BatchNorm2d):
            cur = tape.stage(cur, m, nxt, ACT_NONE, halo=h, halo_kind=hk)
            i += 2
        elif isinstance(nxt, nn.LeakyReLU):
            cur = tape.stage(cur, m, None, ACT_LEAKY, float(nxt.negative_slope), halo=h, halo_kind=hk)
            i += 2
        else:
            cur = tape.stage(cur, m, None, ACT_NONE, halo=h, halo_kind=hk)
            i += 1
    return cur


class G_side(nn.Module):
    """new_multi/networks5_ds.py:623-651."""

    def __init__(self, side_c, df_c, f_size):
        super(G_side, self).__init__()
        self.attention_bs = nn.Sequential(nn.Conv2d(df_c, df_c, 3, 2, padding=1), nn.BatchNorm2d(df_c),
                                          nn.LeakyReLU(0.02), nn.AdaptiveAvgPool2d(1))
        self.at_act = nn.Sigmoid()
        self.side_conv = nn.Sequential(nn.Conv2d(side_c, df_c, 3, 1, padding=1), nn.LeakyReLU(0.02),
                                       nn.BatchNorm2d(df_c), nn.Conv2d(df_c, df_c, 1, 1), nn.LeakyReLU(0.02),
                                       nn.BatchNorm2d(df_c))
        half = int(df_c / 2)
        self.conv = nn.Sequential(nn.Conv2d(df_c, half, 3, 1, padding=1), nn.LeakyReLU(0.02), nn.BatchNorm2d(half),
                                  nn.Conv2d(half, half, 1, 1), nn.LeakyReLU(0.02), nn.BatchNorm2d(half),
                                  nn.UpsamplingBilinear2d(scale_factor=2))

    forward = _parts_only('G_side')


def _g_side(tape, m, s_feature, d_features):
    att = _conv_act_bn_chain(tape, m.attention_bs, d_features)
    s_f = _conv_act_bn_chain(tape, m.side_conv, s_feature)
    fused = tape.gate(d_features, s_f, att, halo=1, halo_kind='zero')          # d + sigmoid(GAP(att)) * s_f
    out = _conv_act_bn_chain(tape, m.conv, fused)
    return tape.bilinear2x(out)


class depth_block(nn.Module):
    """new_multi/networks5_ds.py:655-706."""

    def __init__(self, in_c):
        super(depth_block, self).__init__()
        half = int(in_c / 2)
        self.upconv = nn.ModuleList()
        self.depth_out = nn.ModuleList()
        self.attention_bs = nn.ModuleList()
        for _ in range(4):
            self.upconv.append(nn.Sequential(nn.ConvTranspose2d(in_c, half, 4, 2, padding=1), nn.LeakyReLU(0.02),
                                             nn.BatchNorm2d(half), nn.Conv2d(half, half, 3, 1, padding=1),
                                             nn.LeakyReLU(0.02), nn.BatchNorm2d(half)))
            self.depth_out.append(nn.Sequential(nn.Conv2d(half, 1, 3, 1, padding=1), nn.Tanh()))
            self.attention_bs.append(nn.Sequential(nn.Conv2d(in_c, half, 3, 2, padding=1), nn.LeakyReLU(0.02),
                                                   nn.BatchNorm2d(half), nn.Conv2d(half, half, 3, 2, padding=1),
                                                   nn.LeakyReLU(0.02), nn.AdaptiveAvgPool2d(1)))
        self.at_act = nn.Sigmoid()
        self.conv = nn.Sequential(nn.Conv2d(int(in_c * 2), int(in_c), 3, 1, padding=1), nn.LeakyReLU(0.02),
                                  nn.BatchNorm2d(in_c), nn.Conv2d(int(in_c), half, 3, 1, padding=1),
                                  nn.LeakyReLU(0.02), nn.BatchNorm2d(half))
        self.depconv = nn.Sequential(nn.Conv2d(half, 1, 3, stride=1, padding=1), nn.BatchNorm2d(1))

    forward = _parts_only('depth_block')


def _depth_block(tape, m, in_f):
    n, h, w, _ = in_f.t.shape
    half = m.depth_out[0][0].in_channels
    cat = tape.concat_buffer(n, 2 * h, 2 * w, 4 * half, halo=1, halo_kind='zero')
    outs, slots = [], []
    for i in range(4):
        feat = _conv_act_bn_chain(tape, m.upconv[i], in_f, halo=1, halo_kind='zero')
        o, s = tape.stage(feat, m.depth_out[i][0], None, ACT_TANH, out_nchw=True)
        outs.append(o)
        slots.append(s)
        att = _conv_act_bn_chain(tape, m.attention_bs[i], in_f)
        tape.gate(feat, feat, att, out=cat.slice(i * half, (i + 1) * half))    # sigmoid(at) * f + f
    F = _conv_act_bn_chain(tape, m.conv, cat, halo=1, halo_kind='zero')
    dep1 = _conv_act_bn_chain(tape, m.depconv, F)
    o, s = tape.output_nchw(dep1)
    return outs, slots, o, s


class seg_block(nn.Module):
    """new_multi/networks5_ds.py:708-728."""

    def __init__(self, in_c, out_c):
        super(seg_block, self).__init__()
        self.deconv = nn.Sequential(nn.Conv2d(int(in_c), int(in_c), 3, 1, padding=1), nn.LeakyReLU(0.02),
                                    nn.BatchNorm2d(in_c), nn.Conv2d(in_c, out_c, 1, 1),
                                    nn.UpsamplingBilinear2d(scale_factor=2))
        self.bn = torch.nn.BatchNorm2d(out_c)
        self.lru = nn.LeakyReLU(0.02)

    forward = _parts_only('seg_block')


def _seg_block(tape, m, x):
    v = _conv_act_bn_chain(tape, m.deconv, x)
    v = tape.bilinear2x(v)
    v = tape.norm_act(v, m.bn, ACT_LEAKY, float(m.lru.negative_slope))
    return tape.output_nchw(v)


class R_dep(nn.Module):
    """new_multi/networks5_ds.py:733-821 (the reference's forward also prints a separator line, :791)."""

    def __init__(self):
        super(R_dep, self).__init__()
        self.up0 = DeconvBlock(1024, 512)
        self.AT = nn.ModuleList()
        self.seg_d = seg_block(in_c=128, out_c=28)
        self.dep = depth_block(in_c=128)
        self.AT.append(G_side(side_c=1664, df_c=1024, f_size=40))
        self.AT.append(G_side(side_c=1280, df_c=512, f_size=80))
        self.AT.append(G_side(side_c=512, df_c=256, f_size=160))
        self.dep_out = nn.Conv2d(64, 1, 1, 1)
        self.norm = nn.BatchNorm2d(1)

    def _body(self, tape, s3, s2, s1, d):
        w = tape.input_wants
        v3 = tape.input_nchw(s3, pad=1, pad_kind='zero', want_grad=w[0])
        v2 = tape.input_nchw(s2, pad=1, pad_kind='zero', want_grad=w[1])
        v1 = tape.input_nchw(s1, pad=1, pad_kind='zero', want_grad=w[2])
        vd = tape.input_nchw(d, want_grad=w[3])
        out0 = _g_side(tape, self.AT[0], v3, vd)
        out1 = _g_side(tape, self.AT[1], v2, out0)
        out2 = _g_side(tape, self.AT[2], v1, out1)
        outs, slots = [], []
        for v in (out0, out1, out2):
            o, s = tape.output_nchw(v)
            outs.append(o)
            slots.append(s)
        o, s = _seg_block(tape, self.seg_d, out2)
        outs.append(o)
        slots.append(s)
        dep_o, dep_slots, dep1, dep1_slot = _depth_block(tape, self.dep, out2)
        outs += dep_o + [dep1]
        slots += dep_slots + [dep1_slot]
        return outs, slots, [v3, v2, v1, vd]

    def forward(self, s_features, d_feature):
        outs = graph.run(self, self._body, [s_features[3], s_features[2], s_features[1], d_feature])
        return (outs[0], outs[1], outs[2]), outs[3], (list(outs[4:8]), outs[8])


# ------------------------------------------------------------------------------------------------
# feature discriminator (new_multi/networks5_ds.py:527-566)
# ------------------------------------------------------------------------------------------------
class _Discriminator(nn.Module):
    """PatchGAN with BatchNorm + PReLU; ONE nn.PReLU instance sits at index 1 and 10 of ``model`` (:532,541,559),
    so its parameter appears under both keys and its gradient is the sum of both uses."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm='batch', activation='PReLU'):
        super(_Discriminator, self).__init__()
        norm_layer = get_norm_layer(norm_type=norm)
        nonlinearity = get_nonlinearity_layer(activation_type=activation)
        use_bias = _bias_follows_instance_norm(norm_layer)
        model = [nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=1, bias=use_bias), nonlinearity]
        nf_mult = 1
        for i in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** i, 8)
            model += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=4, stride=2, padding=1, bias=use_bias),
                      norm_layer(ndf * nf_mult), nn.PReLU()]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        model += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=4, stride=1, padding=1, bias=use_bias),
                  norm_layer(ndf * 8), nonlinearity, nn.Conv2d(ndf * nf_mult, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*model)

    def _body(self, tape, x):
        mods = [self.model[j] for j in range(len(self.model))]   # children() would drop the repeated shared PReLU
        v = tape.input_nchw(x, want_grad=tape.input_wants[0])
        cur, i = v, 0
        while i < len(mods):
            m = mods[i]
            if not isinstance(m, nn.Conv2d):
                raise NotImplementedError("unexpected module %s in _Discriminator" % type(m).__name__)
            if i == len(mods) - 1:
                out, slot = tape.stage(cur, m, None, ACT_NONE, out_nchw=True)
                return [out], [slot], [v]
            norm = None
            i += 1
            if isinstance(mods[i], (nn.BatchNorm2d, nn.InstanceNorm2d)):
                norm = mods[i]
                i += 1
            cur = tape.stage(cur, m, norm, ACT_NONE)
            act = mods[i]
            if isinstance(act, nn.PReLU):
                cur = tape.prelu(cur, act)
            elif isinstance(act, nn.LeakyReLU):
                cur = tape.norm_act(cur, None, ACT_LEAKY, float(act.negative_slope))
            elif isinstance(act, nn.ReLU):
                cur = tape.norm_act(cur, None, ACT_RELU)
            else:
                raise NotImplementedError("activation %s" % type(act).__name__)
            i += 1
        raise NotImplementedError("_Discriminator must end with a convolution")

    def forward(self, input):
        return graph.run(self, self._body, [input])[0]


# ------------------------------------------------------------------------------------------------
# losses (new_multi/networks5_ds.py:926-982)
# ------------------------------------------------------------------------------------------------
class GANLoss(nn.Module):
    """new_multi/networks5_ds.py:926-943: always MSE (LSGAN), whatever ``use_lsgan`` says; the label buffers are
    created on the GPU like the reference's (``.cuda()``)."""

    def __init__(self, use_lsgan=False, target_real_label=1.0, target_fake_label=0.0):
        super(GANLoss, self).__init__()
        self.register_buffer('real_label', torch.tensor(target_real_label).cuda())
        self.register_buffer('fake_label', torch.tensor(target_fake_label).cuda())
        self._labels = (float(target_real_label), float(target_fake_label))

    def get_target_tensor(self, input, target_is_real):
        return (self.real_label if target_is_real else self.fake_label).expand_as(input)

    def __call__(self, input, target_is_real):
        return losses.mse_const(input, self._labels[0] if target_is_real else self._labels[1])


class BCEDepLoss(nn.Module):
    """new_multi/networks5_ds.py:947-956: BCE((x+1)/2*o_m, (t+1)/2*o_m) + BCE((x+1)/2*z_m, (t+1)/2*z_m) + 50*L1(x, t).
    With the masks ``get_masks(target)`` produces (the only way the reference calls it, model5.py:526-536) this is
    one fused kernel; other masks raise."""

    def __init__(self):
        super(BCEDepLoss, self).__init__()

    def __call__(self, input, target, o_m, z_m):
        return losses.bcedep(input, target, o_m, z_m)


def get_masks(target):
    """new_multi/networks5_ds.py:973-982: o_m = (target == 1), z_m = (target == -1) as float tensors."""
    return losses.depth_bin_masks(target)
