"""Drop-in ``models/seg_network.py`` (second half): the two-headed U-Net generator
``_UNetGenerator`` that ``models/seg_model.py:8`` builds (``define_G(..., model_type='UNet')``), on the graph engine.

One encoder (c7s1-ngf + three double-conv blocks, AvgPool2 after each), ``7 - layers`` inception blocks and the centre
up-block, then TWO decoders — ``*_syn`` with 22 output classes and ``*_real`` with 28 — of which ``forward(input,
syn_or_real)`` runs one (models/seg_network.py:441-559).  Same constructor signature, module tree and ``state_dict``
keys as the reference, including its single shared ``nn.PReLU``.  ``forward`` returns ``[center_in, output1]``.

The blocks are those of ``models/encoder_decoder.py`` (the reference repeats their definitions, :155-285) and run
through the same tape code (``encoder_decoder._EncoderBlock.run`` ...); encoder, centre and the selected decoder form ONE
tape: the skips stay bf16 NHWC values (no fp32 round trip between an encoder and a decoder module as in SegCycle), and
every ``torch.cat`` is a preallocated buffer whose channel slices the producers write.

``define_G`` mirrors models/seg_network.py:112-124 for ``model_type='UNet'``; the other generator variants of the file
(``_ResGenerator``, ``_PreUNet16`` with its torchvision VGG16 encoder) raise ``NotImplementedError``.
``_Discriminator`` / ``_MultiscaleDiscriminator`` / ``define_D`` (:126-132, 561-627): below; one scale (num_D = 1, the default).
"""
import torch
import torch.nn as nn

from . import graph
from .encoder_decoder import (_DecoderUpBlock, _EncoderBlock, _InceptionBlock, _OutputBlock, _activate, _uses_bias,
                              get_nonlinearity_layer, get_norm_layer)
from .ops import ACT_NONE, ACT_TANH

NC_SYN = 22      # models/seg_network.py:483-484
NC_REAL = 28


class _UNetGenerator(nn.Module):
    """models/seg_network.py:441-559."""

    def __init__(self, input_nc, output_nc, ngf=64, layers=4, norm='batch', activation='PReLU', drop_rate=0,
                 add_noise=False, gpu_ids=[0], weight=0.1):
        super().__init__()
        if layers != 4:
            # the reference registers the extra `down<i>` encoders for layers > 4 but its forward never feeds their skips
            # back (the `up<i>` blocks are commented out, :486-488): only the default depth is mirrored
            raise NotImplementedError("_UNetGenerator with layers != 4 on the B200 path")
        if add_noise:
            raise NotImplementedError("GaussianNoiseLayer (add_noise=True) on the B200 path")
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
        center = [_InceptionBlock(ngf * 8, ngf * 8, norm_layer, nonlinearity, 7 - layers, drop_rate, use_bias)
                  for _ in range(7 - layers)]
        center += [_DecoderUpBlock(ngf * 8, ngf * 8, ngf * 4, norm_layer, nonlinearity, use_bias)]
        self.center = nn.Sequential(*center)
        for head, nc in (('syn', NC_SYN), ('real', NC_REAL)):
            setattr(self, 'deconv4_' + head, _DecoderUpBlock(ngf * (4 + 4), ngf * 8, ngf * 2, norm_layer, nonlinearity, use_bias))
            setattr(self, 'deconv3_' + head, _DecoderUpBlock(ngf * (2 + 2) + nc, ngf * 4, ngf, norm_layer, nonlinearity, use_bias))
            setattr(self, 'deconv2_' + head, _DecoderUpBlock(ngf * (1 + 1) + nc, ngf * 2, int(ngf / 2), norm_layer, nonlinearity,
                                                             use_bias))
            setattr(self, 'output4_' + head, _OutputBlock(ngf * (4 + 4), nc, 3, use_bias))
            setattr(self, 'output3_' + head, _OutputBlock(ngf * (2 + 2) + nc, nc, 3, use_bias))
            setattr(self, 'output2_' + head, _OutputBlock(ngf * (1 + 1) + nc, nc, 3, use_bias))
            setattr(self, 'output1_' + head, _OutputBlock(int(ngf / 2) + nc, nc, 7, use_bias))
        self.upsample = nn.Upsample(scale_factor=2, mode='nearest')
        # the reference builds the modules in the order deconv4/3/2_syn, output4..1_syn, deconv4/3/2_real, output4..1_real
        order = ['pool', 'conv1', 'conv2', 'conv3', 'conv4', 'center']
        for head in ('syn', 'real'):
            order += ['deconv4_' + head, 'deconv3_' + head, 'deconv2_' + head, 'output4_' + head, 'output3_' + head,
                      'output2_' + head, 'output1_' + head]
        order += ['upsample']
        self._modules = type(self._modules)((k, self._modules[k]) for k in order)

    def _body(self, tape, x, head):
        c1 = self.conv1
        img = tape.input_nchw(x, pad=3, pad_kind='reflect', first_conv=c1[1], want_grad=tape.input_wants[0])
        v = tape.stage(img, c1[1], c1[2], ACT_NONE, reflect=3)
        v = _activate(tape, v, c1[3])
        conv1 = tape.avgpool2(v, halo=1, halo_kind='zero')
        conv2 = tape.avgpool2(self.conv2.run(tape, conv1), halo=1, halo_kind='zero')
        conv3 = tape.avgpool2(self.conv3.run(tape, conv2), halo=1, halo_kind='zero')
        center_in = tape.avgpool2(self.conv4.run(tape, conv3))
        cur = center_in
        blocks = list(self.center)
        for blk in blocks[:-1]:
            cur = blk.run(tape, cur)
        up_center = blocks[-1]
        nc = NC_SYN if head == 'syn' else NC_REAL
        n = x.shape[0]
        g = lambda name: getattr(self, name + '_' + head)  # noqa: E731

        def level(skip, c_up, scale, with_prev_output, like):
            """Concatenation buffer [up-block output, scaled skip, upsampled previous output] (:527-537)."""
            c_skip = skip.c if skip is not None else 0
            h, w = (skip.t.shape[1], skip.t.shape[2]) if skip is not None else (2 * like.t.shape[1], 2 * like.t.shape[2])
            total = c_up + c_skip + (nc if with_prev_output else 0)
            cat = tape.concat_buffer(n, h, w, total)
            if skip is not None:
                tape.scale(skip, scale, out=cat.slice(c_up, c_up + c_skip))
            return cat, cat.slice(0, c_up), (cat.slice(c_up + c_skip, total) if with_prev_output else None)

        def output(block, xp, into):
            conv = block.model[1]
            raw = tape.stage(xp, conv, None, ACT_NONE, reflect=conv.kernel_size[0] // 2)
            tape.nearest2x(tape.tanh(raw), out=into)

        w0 = self.weight
        curp = tape.norm_act(cur, None, ACT_NONE, halo=1, halo_kind='reflect')      # nn.ReflectionPad2d(1) of the up-block
        cat4, up4, _ = level(conv3, up_center.model[4].out_channels, w0, False, None)
        up_center.run(tape, curp, up4)
        cat4p = tape.norm_act(cat4, None, ACT_NONE, halo=1, halo_kind='reflect')

        cat3, up3, prev3 = level(conv2, g('deconv4').model[4].out_channels, w0 * 0.5, True, None)
        g('deconv4').run(tape, cat4p, up3)
        output(g('output4'), cat4p, prev3)
        cat3p = tape.norm_act(cat3, None, ACT_NONE, halo=1, halo_kind='reflect')

        cat2, up2, prev2 = level(conv1, g('deconv3').model[4].out_channels, w0 * 0.1, True, None)
        g('deconv3').run(tape, cat3p, up2)
        output(g('output3'), cat3p, prev2)
        cat2p = tape.norm_act(cat2, None, ACT_NONE, halo=1, halo_kind='reflect')

        cat1, up1, prev1 = level(None, g('deconv2').model[4].out_channels, 0.0, True, conv1)
        g('deconv2').run(tape, cat2p, up1)
        output(g('output2'), cat2p, prev1)
        k1 = g('output1').model[1].kernel_size[0]
        cat1p = tape.norm_act(cat1, None, ACT_NONE, halo=k1 // 2, halo_kind='reflect')
        out1, s1 = tape.stage(cat1p, g('output1').model[1], None, ACT_TANH, reflect=k1 // 2, out_nchw=True)
        cin, s0 = tape.output_nchw(center_in)
        return [cin, out1], [s0, s1], [img]

    def forward(self, input, syn_or_real):
        head = 'syn' if syn_or_real == 'syn' else 'real'       # the reference's else-branch takes every other value
        return list(graph.run(self, lambda tape, x: self._body(tape, x, head), [input]))


class _Discriminator(nn.Module):
    """models/seg_network.py:585-627: PatchGAN with 4x4 convolutions, BatchNorm and ONE shared non-linearity instance at
    every activation slot (``nonlinearity`` is created once, :592): with PReLU its single slope appears under four keys
    and receives the sum of the four uses' gradients.  Runs through the tape body of the seg/depth feature discriminator
    (networks5_ds._Discriminator, the same layer pattern with per-layer PReLUs)."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm='batch', activation='PReLU', gpu_ids=[]):
        super().__init__()
        self.gpu_ids = gpu_ids
        norm_layer = get_norm_layer(norm_type=norm)
        if norm_layer is None:
            raise NotImplementedError("norm='none' (the reference itself fails on norm_layer(ndf))")
        nonlinearity = get_nonlinearity_layer(activation_type=activation)
        use_bias = _uses_bias(norm_layer)
        if n_layers != 3 and min(2 ** n_layers, 8) != 8:
            # the reference normalises the last hidden layer with norm_layer(ndf * 8) whatever its width is (:615)
            raise NotImplementedError("_Discriminator with n_layers < 3 (the reference's BatchNorm width does not match)")
        model = [nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=1, bias=use_bias), nonlinearity]
        nf_mult = 1
        for i in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** i, 8)
            model += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=4, stride=2, padding=1, bias=use_bias),
                      norm_layer(ndf * nf_mult), nonlinearity]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        model += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=4, stride=1, padding=1, bias=use_bias),
                  norm_layer(ndf * 8), nonlinearity, nn.Conv2d(ndf * nf_mult, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*model)

    def forward(self, input):
        from .networks5_ds import _Discriminator as _TapeBody
        return graph.run(self, lambda tape, x: _TapeBody._body(self, tape, x), [input])[0]


class _MultiscaleDiscriminator(nn.Module):
    """models/seg_network.py:561-583: ``num_D`` discriminators over an image pyramid; returns the list of their outputs.
    ``define_D`` builds it with num_D = 1 (the default of models/seg_model.py), which never touches the 3x3 stride-2
    ``count_include_pad=False`` average pooling between the scales; num_D > 1 is refused."""

    def __init__(self, input_nc, ndf=64, n_layers=3, num_D=1, norm='batch', activation='PReLU', gpu_ids=[]):
        super().__init__()
        if num_D != 1:
            raise NotImplementedError("_MultiscaleDiscriminator with num_D > 1 (image-pyramid pooling) on the B200 path")
        self.num_D = num_D
        self.gpu_ids = gpu_ids
        for i in range(num_D):
            setattr(self, 'scale' + str(i), _Discriminator(input_nc, ndf, n_layers, norm, activation, gpu_ids))
        self.downsample = nn.AvgPool2d(kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)

    def forward(self, input):
        return [getattr(self, 'scale' + str(i))(input) for i in range(self.num_D)]


def define_D(input_nc, ndf=64, n_layers=3, num_D=1, norm='batch', activation='PReLU', init_type='xavier', gpu_ids=[]):
    """models/seg_network.py:126-132."""
    net = _MultiscaleDiscriminator(input_nc, ndf, n_layers, num_D, norm, activation, gpu_ids)
    from .networks import init_weights
    init_weights(net, init_type)
    if len(gpu_ids) > 0:
        net.to("cuda:%d" % gpu_ids[0])
    return net


def define_G(input_nc, output_nc, ngf=64, layers=4, norm='batch', activation='PReLU', model_type='UNet',
             init_type='xavier', drop_rate=0, add_noise=False, gpu_ids=[], weight=0.1):
    """models/seg_network.py:112-124 (init_net: xavier / normal initialisation, then .to(gpu_ids[0]))."""
    if model_type != 'UNet':
        raise NotImplementedError("model_type %r: only the U-Net generator of models/seg_model.py is on the B200 path"
                                  % (model_type,))
    net = _UNetGenerator(input_nc, output_nc, ngf, layers, norm, activation, drop_rate, add_noise, gpu_ids, weight)
    from .networks import init_weights
    init_weights(net, init_type)
    if len(gpu_ids) > 0:
        net.to("cuda:%d" % gpu_ids[0])
    return net
