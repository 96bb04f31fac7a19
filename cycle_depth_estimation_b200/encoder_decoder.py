"""Drop-in ``models/encoder_decoder.py``: the U-Net task network of SegCycle
(``models/seg_cycle.py:44-57``) — ``_UNetEncoder`` / ``_UNetDecoder`` and their blocks — on the graph engine.

Same class names, constructor signatures, module tree and therefore ``state_dict`` keys as the reference
(``models/encoder_decoder.py:30-208``), including its quirk that ONE ``nn.PReLU`` instance is shared by every block
of a network (``get_nonlinearity_layer`` is called once per encoder / decoder, :134,:175), so the single slope
parameter appears under many keys and receives the sum of all its uses' gradients.

Dataflow on the tape (``graph.py``), all bf16 NHWC, no ``torch.cat``:
* encoder: c7s1-64 (row-packed image operand) -> BN -> PReLU -> AvgPool2, three (conv3x3-BN-PReLU)x2 + AvgPool2
  blocks, then ``7 - layers`` inception blocks whose three dilated branches write channel slices of one buffer;
* decoder: every ``torch.cat([...], 1)`` (:196-205) is a preallocated buffer whose slices are written by the
  producers — the PReLU of the previous up-block, the scaled skip (``cdb_scale``) and the nearest-neighbour
  upsampling of the previous output block (``cdb_nearest2x_fwd``); the buffer is reflect-padded once and read by
  both consumers (the up-block and the output block), whose two halo-covering gradients ``Tape.gather`` folds.
"""
import functools

import torch
import torch.nn as nn

from . import graph
from .ops import ACT_NONE, ACT_TANH


def get_norm_layer(norm_type='batch'):
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False)
    if norm_type == 'none':
        return None
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def get_nonlinearity_layer(activation_type='PReLU'):
    if activation_type == 'ReLU':
        return nn.ReLU(True)
    if activation_type == 'SELU':
        return nn.SELU(True)
    if activation_type == 'LeakyReLU':
        return nn.LeakyReLU(0.1, True)
    if activation_type == 'PReLU':
        return nn.PReLU()
    raise NotImplementedError('activation layer [%s] is not found' % activation_type)


def _uses_bias(norm_layer):
    func = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    return func == nn.InstanceNorm2d


def _parts_only(self, *a, **k):
    raise RuntimeError("%s is a parameter container of the B200 path; call the enclosing _UNetEncoder / "
                       "_UNetDecoder" % type(self).__name__)


def _activate(tape, v, act, halo=0, halo_kind=None, out=None):
    """The network's non-linearity on the tape (PReLU with its device-resident slope, or a fixed-slope kind)."""
    if isinstance(act, nn.PReLU):
        return tape.prelu(v, act, halo=halo, halo_kind=halo_kind, out=out)
    if isinstance(act, nn.LeakyReLU):
        return tape.norm_act(v, None, graph.ACT_LEAKY, float(act.negative_slope), halo=halo, halo_kind=halo_kind, out=out)
    if isinstance(act, nn.ReLU):
        return tape.norm_act(v, None, graph.ACT_RELU, halo=halo, halo_kind=halo_kind, out=out)
    raise NotImplementedError("activation %s on the B200 path" % type(act).__name__)


class _EncoderBlock(nn.Module):
    """conv3x3 - norm - act - conv3x3 - norm - act (models/encoder_decoder.py:30-46)."""

    def __init__(self, input_nc, middle_nc, output_nc, norm_layer=nn.BatchNorm2d, nonlinearity=nn.PReLU(), use_bias=False):
        super().__init__()
        self.model = nn.Sequential(
            nn.Conv2d(input_nc, middle_nc, kernel_size=3, stride=1, padding=1, bias=use_bias), norm_layer(middle_nc),
            nonlinearity,
            nn.Conv2d(middle_nc, output_nc, kernel_size=3, stride=1, padding=1, bias=use_bias), norm_layer(output_nc),
            nonlinearity)

    forward = _parts_only

    def run(self, tape, x):
        m = self.model
        a = tape.stage(x, m[0], m[1], ACT_NONE)
        a = _activate(tape, a, m[2], halo=1, halo_kind='zero')
        b = tape.stage(a, m[3], m[4], ACT_NONE)
        return _activate(tape, b, m[5])


class _InceptionBlock(nn.Module):
    """``width`` dilated 3x3 branches (reflect pad = dilation = 2i+1) -> cat -> norm -> act -> reflect-pad conv3x3 ->
    norm -> (+ x) -> act (models/encoder_decoder.py:47-82)."""

    def __init__(self, input_nc, output_nc, norm_layer=nn.BatchNorm2d, nonlinearity=nn.PReLU(), width=1, drop_rate=0,
                 use_bias=False):
        super().__init__()
        self.width = width
        self.drop_rate = drop_rate
        for i in range(width):
            setattr(self, 'layer' + str(i), nn.Sequential(
                nn.ReflectionPad2d(i * 2 + 1),
                nn.Conv2d(input_nc, output_nc, kernel_size=3, padding=0, dilation=i * 2 + 1, bias=use_bias)))
        self.norm1 = norm_layer(output_nc * width)
        self.norm2 = norm_layer(output_nc)
        self.nonlinearity = nonlinearity
        self.branch1x1 = nn.Sequential(
            nn.ReflectionPad2d(1),
            nn.Conv2d(output_nc * width, output_nc, kernel_size=3, padding=0, bias=use_bias))

    forward = _parts_only

    def run(self, tape, x):
        n, h, w, _ = x.t.shape
        co = self.norm2.num_features
        cat = tape.concat_buffer(n, h, w, co * self.width)
        for i in range(self.width):
            pad = i * 2 + 1
            if pad >= h or pad >= w:
                raise RuntimeError("reflection padding %d needs a feature map larger than %dx%d" % (pad, h, w))
            xp = tape.norm_act(x, None, ACT_NONE, halo=pad, halo_kind='reflect')      # nn.ReflectionPad2d(pad)
            tape.stage(xp, getattr(self, 'layer' + str(i))[1], None, ACT_NONE, reflect=pad,
                       out=cat.slice(i * co, (i + 1) * co))
        u = tape.norm_act(cat, self.norm1, ACT_NONE)
        u = _activate(tape, u, self.nonlinearity)
        up = tape.norm_act(u, None, ACT_NONE, halo=1, halo_kind='reflect')
        v = tape.stage(up, self.branch1x1[1], self.norm2, ACT_NONE, reflect=1, res=x)
        if self.drop_rate > 0:
            v = tape.dropout(v, self.drop_rate)
        return _activate(tape, v, self.nonlinearity)


class _DecoderUpBlock(nn.Module):
    """reflect-pad conv3x3 - norm - act - ConvTranspose 3x3 s2 - norm - act (models/encoder_decoder.py:84-101)."""

    def __init__(self, input_nc, middle_nc, output_nc, norm_layer=nn.BatchNorm2d, nonlinearity=nn.PReLU(), use_bias=False):
        super().__init__()
        self.model = nn.Sequential(
            nn.ReflectionPad2d(1),
            nn.Conv2d(input_nc, middle_nc, kernel_size=3, stride=1, padding=0, bias=use_bias), norm_layer(middle_nc),
            nonlinearity,
            nn.ConvTranspose2d(middle_nc, output_nc, kernel_size=3, stride=2, padding=1, output_padding=1),
            norm_layer(output_nc), nonlinearity)

    forward = _parts_only

    def run(self, tape, xp, out):
        """xp: input with a materialised reflect halo of 1; out: the channel slice that receives the result."""
        m = self.model
        a = tape.stage(xp, m[1], m[2], ACT_NONE, reflect=1)
        a = _activate(tape, a, m[3])
        b = tape.stage(a, m[4], m[5], ACT_NONE)
        return _activate(tape, b, m[6], out=out)


class _OutputBlock(nn.Module):
    """reflect-pad conv kxk - tanh (models/encoder_decoder.py:103-117)."""

    def __init__(self, input_nc, output_nc, kernel_size=3, use_bias=False):
        super().__init__()
        self.model = nn.Sequential(
            nn.ReflectionPad2d(int(kernel_size / 2)),
            nn.Conv2d(input_nc, output_nc, kernel_size=kernel_size, padding=0, bias=use_bias), nn.Tanh())

    forward = _parts_only


class _UNetEncoder(nn.Module):
    """models/encoder_decoder.py:122-162. forward(input) -> [conv1, conv2, conv3, center_in, center_out] (fp32 NCHW)."""

    def __init__(self, input_nc, ngf=64, layers=4, norm='batch', activation='PReLU', drop_rate=0, weight=0.1):
        super().__init__()
        self.layers = layers
        self.weight = weight
        norm_layer = get_norm_layer(norm_type=norm)
        if norm_layer is None:
            raise NotImplementedError("norm='none' (the reference itself fails on norm_layer(ngf))")
        nonlinearity = get_nonlinearity_layer(activation_type=activation)
        use_bias = _uses_bias(norm_layer)
        self.pool = nn.AvgPool2d(kernel_size=2, stride=2)
        self.conv1 = nn.Sequential(nn.ReflectionPad2d(3),
                                   nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias), norm_layer(ngf),
                                   nonlinearity)
        self.conv2 = _EncoderBlock(ngf, ngf * 2, ngf * 2, norm_layer, nonlinearity, use_bias)
        self.conv3 = _EncoderBlock(ngf * 2, ngf * 4, ngf * 4, norm_layer, nonlinearity, use_bias)
        self.conv4 = _EncoderBlock(ngf * 4, ngf * 8, ngf * 8, norm_layer, nonlinearity, use_bias)
        self.center = nn.Sequential(*[
            _InceptionBlock(ngf * 8, ngf * 8, norm_layer, nonlinearity, 7 - layers, drop_rate, use_bias)
            for _ in range(7 - layers)])

    def _body(self, tape, x):
        c1 = self.conv1
        img = tape.input_nchw(x, pad=3, pad_kind='reflect', first_conv=c1[1], want_grad=tape.input_wants[0])
        v = tape.stage(img, c1[1], c1[2], ACT_NONE, reflect=3)
        v = _activate(tape, v, c1[3])
        conv1 = tape.avgpool2(v, halo=1, halo_kind='zero')
        conv2 = tape.avgpool2(self.conv2.run(tape, conv1), halo=1, halo_kind='zero')
        conv3 = tape.avgpool2(self.conv3.run(tape, conv2), halo=1, halo_kind='zero')
        center_in = tape.avgpool2(self.conv4.run(tape, conv3))
        cur = center_in
        for blk in self.center:
            cur = blk.run(tape, cur)
        outs, slots = [], []
        for v in (conv1, conv2, conv3, center_in, cur):
            o, s = tape.output_nchw(v)
            outs.append(o)
            slots.append(s)
        return outs, slots, [img]

    def forward(self, input):
        return list(graph.run(self, self._body, [input]))


class _UNetDecoder(nn.Module):
    """models/encoder_decoder.py:164-208. forward([conv1, conv2, conv3, center_in, center_out]) ->
    [center_in, output4, output3, output2, output1]."""

    def __init__(self, output_nc, ngf=64, layers=4, norm='batch', activation='PReLU', weight=0.1):
        super().__init__()
        self.layers = layers
        self.weight = weight
        norm_layer = get_norm_layer(norm_type=norm)
        if norm_layer is None:
            raise NotImplementedError("norm='none' (the reference itself fails on norm_layer(ngf))")
        nonlinearity = get_nonlinearity_layer(activation_type=activation)
        use_bias = _uses_bias(norm_layer)
        self.pool = nn.AvgPool2d(kernel_size=2, stride=2)
        self.deconv_center = _DecoderUpBlock(ngf * 8, ngf * 8, ngf * 4, norm_layer, nonlinearity, use_bias)
        self.deconv4 = _DecoderUpBlock(ngf * (4 + 4), ngf * 8, ngf * 2, norm_layer, nonlinearity, use_bias)
        self.deconv3 = _DecoderUpBlock(ngf * (2 + 2) + output_nc, ngf * 4, ngf, norm_layer, nonlinearity, use_bias)
        self.deconv2 = _DecoderUpBlock(ngf * (1 + 1) + output_nc, ngf * 2, int(ngf / 2), norm_layer, nonlinearity,
                                       use_bias)
        self.output4 = _OutputBlock(ngf * (4 + 4), output_nc, 3, use_bias)
        self.output3 = _OutputBlock(ngf * (2 + 2) + output_nc, output_nc, 3, use_bias)
        self.output2 = _OutputBlock(ngf * (1 + 1) + output_nc, output_nc, 3, use_bias)
        self.output1 = _OutputBlock(int(ngf / 2) + output_nc, output_nc, 7, use_bias)
        self.upsample = nn.Upsample(scale_factor=2, mode='nearest')

    def _output(self, tape, block, xp, into):
        """Intermediate output block: tanh(conv(reflect-padded input)); returns the fp32 tensor + slot and writes the
        nearest-neighbour upsampling into the next level's concatenation slice."""
        conv = block.model[1]
        raw = tape.stage(xp, conv, None, ACT_NONE, reflect=conv.kernel_size[0] // 2)
        o = tape.tanh(raw)
        tape.nearest2x(o, out=into)
        return tape.output_nchw(o)

    def _body(self, tape, conv1, conv2, conv3, center_out):
        wants = tape.input_wants
        v1 = tape.input_nchw(conv1, want_grad=wants[0])
        v2 = tape.input_nchw(conv2, want_grad=wants[1])
        v3 = tape.input_nchw(conv3, want_grad=wants[2])
        vc = tape.input_nchw(center_out, pad=1, pad_kind='reflect', want_grad=wants[3])
        nc = self.output1.model[1].out_channels
        n = conv1.shape[0]

        def level(skip, up_block, c_up, scale, with_prev_output):
            """Concatenation buffer [up-block output, scaled skip, upsampled previous output] at the skip's size."""
            c_skip = skip.c if skip is not None else 0
            h, w = (skip.t.shape[1], skip.t.shape[2]) if skip is not None else (2 * v1.t.shape[1], 2 * v1.t.shape[2])
            total = c_up + c_skip + (nc if with_prev_output else 0)
            cat = tape.concat_buffer(n, h, w, total)
            if skip is not None:
                tape.scale(skip, scale, out=cat.slice(c_up, c_up + c_skip))
            prev_slot = cat.slice(c_up + c_skip, total) if with_prev_output else None
            return cat, cat.slice(0, c_up), prev_slot

        w0 = self.weight
        c4 = self.deconv_center.model[4].out_channels
        cat4, up4, _ = level(v3, self.deconv_center, c4, w0, False)
        self.deconv_center.run(tape, vc, up4)
        cat4p = tape.norm_act(cat4, None, ACT_NONE, halo=1, halo_kind='reflect')

        c3 = self.deconv4.model[4].out_channels
        cat3, up3, prev3 = level(v2, self.deconv4, c3, w0 * 0.5, True)
        self.deconv4.run(tape, cat4p, up3)
        out4, s4 = self._output(tape, self.output4, cat4p, prev3)
        cat3p = tape.norm_act(cat3, None, ACT_NONE, halo=1, halo_kind='reflect')

        c2 = self.deconv3.model[4].out_channels
        cat2, up2, prev2 = level(v1, self.deconv3, c2, w0 * 0.1, True)
        self.deconv3.run(tape, cat3p, up2)
        out3, s3 = self._output(tape, self.output3, cat3p, prev2)
        cat2p = tape.norm_act(cat2, None, ACT_NONE, halo=1, halo_kind='reflect')

        c1 = self.deconv2.model[4].out_channels
        cat1, up1, prev1 = level(None, self.deconv2, c1, 0.0, True)
        self.deconv2.run(tape, cat2p, up1)
        out2, s2 = self._output(tape, self.output2, cat2p, prev1)
        k1 = self.output1.model[1].kernel_size[0]
        cat1p = tape.norm_act(cat1, None, ACT_NONE, halo=k1 // 2, halo_kind='reflect')
        out1, s1 = tape.stage(cat1p, self.output1.model[1], None, ACT_TANH, reflect=k1 // 2, out_nchw=True)
        return [out4, out3, out2, out1], [s4, s3, s2, s1], [v1, v2, v3, vc]

    def forward(self, input):
        conv1, conv2, conv3, center_in, center_out = input
        outs = graph.run(self, self._body, [conv1, conv2, conv3, center_out])
        return [center_in] + list(outs)
