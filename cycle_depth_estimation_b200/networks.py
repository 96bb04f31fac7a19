"""Drop-in replacement for the reference's ``models/networks.py``.

Same public names, constructor signatures, module tree and ``state_dict`` keys as the reference
(models/networks.py:12-389), so ``cycle_gan_model.py`` / ``pix2pix_model.py`` / ``test_model.py`` can
import this module instead.  The leaf modules are the stock ``torch.nn`` classes used purely as
parameter holders; ``forward`` of every network runs the fused B200 engine (``engine.py``) — no
cuDNN, no CPU path.
"""
import functools

import torch
import torch.nn as nn
from torch.nn import init
from torch.optim import lr_scheduler

from . import engine, graph, losses, ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_TANH
from .ops import get_precision, precision, set_precision  # noqa: F401  (re-exported: the network precision switch)


# ------------------------------------------------------------------------------------------------
# helpers (models/networks.py:12-70)
# ------------------------------------------------------------------------------------------------
def get_norm_layer(norm_type='instance'):
    """models/networks.py:12-22."""
    table = {
        'batch': lambda: functools.partial(nn.BatchNorm2d, affine=True),
        'instance': lambda: functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False),
        'none': lambda: None,
    }
    if norm_type not in table:
        raise NotImplementedError('normalization layer [%s] is not found' % norm_type)
    return table[norm_type]()


def get_scheduler(optimizer, opt):
    """models/networks.py:24-38 (the lambda rule is hard-wired to 1 - max(0, epoch-10)/30 there)."""
    policy = opt.lr_policy
    if policy == 'lambda':
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda epoch: 1.0 - max(0, epoch - 10) / float(30))
    if policy == 'step':
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if policy == 'plateau':
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.2, threshold=0.01, patience=5)
    if policy == 'cosine':
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.niter, eta_min=0)
    # the reference RETURNS (does not raise) the exception object here (models/networks.py:37)
    return NotImplementedError('learning rate policy [%s] is not implemented', policy)


def init_weights(net, init_type='normal', gain=0.02):
    """models/networks.py:40-61: N(0, gain) (or xavier/kaiming/orthogonal) on Conv*/Linear weights,
    zero biases, BatchNorm2d weight ~ N(1, gain)."""
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
    engine.invalidate_packed_weights()  # the fills above write through .data


def init_net(net, init_type='normal', init_gain=0.02, gpu_ids=[]):
    """models/networks.py:63-70: moves to gpu_ids[0] (IndexError on an empty list, as there)."""
    net.to(gpu_ids[0])
    init_weights(net, init_type, gain=init_gain)
    return net


def define_G(input_nc, output_nc, ngf, netG, norm='batch', use_dropout=False, init_type='normal',
             init_gain=0.02, gpu_ids=[]):
    """models/networks.py:73-91."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netG in ('resnet_9blocks', 'resnet_6blocks'):
        net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                              n_blocks=9 if netG == 'resnet_9blocks' else 6)
    elif netG in ('unet_128', 'unet_256'):
        net = UnetGenerator(input_nc, output_nc, 7 if netG == 'unet_128' else 8, ngf, norm_layer=norm_layer,
                            use_dropout=use_dropout)
    else:
        raise NotImplementedError('Generator model name [%s] is not recognized' % netG)
    return init_net(net, init_type, init_gain, gpu_ids)


def define_D(input_nc, ndf, netD, n_layers_D=3, norm='batch', use_sigmoid=False, init_type='normal',
             init_gain=0.02, gpu_ids=[]):
    """models/networks.py:94-107 ('basic' ignores n_layers_D there, and so does this)."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netD == 'basic':
        net = NLayerDiscriminator(input_nc, ndf, n_layers=3, norm_layer=norm_layer, use_sigmoid=use_sigmoid)
    elif netD == 'pixel':
        net = PixelDiscriminator(input_nc, ndf, norm_layer=norm_layer, use_sigmoid=use_sigmoid)
    else:
        raise NotImplementedError('Discriminator model name [%s] is not recognized' % netD)
    return init_net(net, init_type, init_gain, gpu_ids)


def _bias_follows_instance_norm(norm_layer):
    """use_bias rule of models/networks.py:152-155."""
    target = norm_layer.func if type(norm_layer) == functools.partial else norm_layer
    return target == nn.InstanceNorm2d


# ------------------------------------------------------------------------------------------------
# GANLoss (models/networks.py:119-138)
# ------------------------------------------------------------------------------------------------
class GANLoss(nn.Module):
    """LSGAN (MSE) or vanilla (BCE) loss against a constant label. The label buffers stay in the
    state_dict as in the reference; the loss itself is one fused kernel (forward value + gradient)."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0):
        super(GANLoss, self).__init__()
        self.register_buffer('real_label', torch.tensor(target_real_label))
        self.register_buffer('fake_label', torch.tensor(target_fake_label))
        self.use_lsgan = use_lsgan
        self._labels = (float(target_real_label), float(target_fake_label))

    def get_target_tensor(self, input, target_is_real):
        return (self.real_label if target_is_real else self.fake_label).expand_as(input)

    def __call__(self, input, target_is_real):
        label = self._labels[0] if target_is_real else self._labels[1]
        if self.use_lsgan:
            return losses.mse_const(input, label)
        return losses.bce_const(input, label)


# ------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------
class _FusedNet(nn.Module):
    """Base of the drop-in networks: compiles `self._body()` once and runs it on the engine."""

    def _body(self):
        raise NotImplementedError

    def _plan(self):
        plan = self.__dict__.get('_cdb_plan')
        if plan is None:
            stages, final = engine.compile_chain(list(self._body().children()))
            plan = engine.Plan(stages, final)
            self.__dict__['_cdb_plan'] = plan
        return plan

    def _chain_body(self, tape, x):
        """The compiled stage list executed on the general tape (graph.py) — the route of the fp32-storage
        precisions ('tf32', 'tf32x3'), which the bf16-specialised chain engine does not implement."""
        plan = self._plan()
        st0 = plan.stages[0]
        v0 = tape.input_nchw(x, pad=st0.reflect, pad_kind='reflect' if st0.reflect else None, first_conv=st0.conv,
                             want_grad=tape.input_wants[0])
        vals = {0: v0}
        out = slot = None
        for idx, st in enumerate(plan.stages):
            src = vals[st.src]
            if idx == len(plan.stages) - 1:
                out, slot = tape.stage(src, st.conv, None, st.act, st.slope, reflect=st.reflect, out_nchw=True)
                break
            halo, zh = plan.halo[st.dst], plan.zero_halo[st.dst]
            vals[st.dst] = tape.stage(src, st.conv, st.norm, st.act, st.slope,
                                      res=vals[st.res] if st.res is not None else None, reflect=st.reflect,
                                      halo=halo or zh, halo_kind='reflect' if halo else ('zero' if zh else None))
        return [out], [slot], [v0]

    def forward(self, input):
        if ops.get_precision() != 'bf16' or getattr(self, '_cdb_force_tape', False):
            return graph.run(self, self._chain_body, [input])[0]
        return engine.run_network(self, self._plan(), input)


class ResnetGenerator(_FusedNet):
    """models/networks.py:145-191: c7s1-ngf, two stride-2 downsamplings, n_blocks residual blocks, two
    transposed-conv upsamplings, c7s1-output_nc, tanh."""

    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6,
                 padding_type='reflect'):
        assert (n_blocks >= 0)
        super(ResnetGenerator, self).__init__()
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        use_bias = _bias_follows_instance_norm(norm_layer)
        layers = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias),
                  norm_layer(ngf), nn.ReLU(True)]
        ch = ngf
        for _ in range(2):  # downsampling
            layers += [nn.Conv2d(ch, ch * 2, kernel_size=3, stride=2, padding=1, bias=use_bias),
                       norm_layer(ch * 2), nn.ReLU(True)]
            ch *= 2
        layers += [ResnetBlock(ch, padding_type=padding_type, norm_layer=norm_layer, use_dropout=use_dropout,
                               use_bias=use_bias) for _ in range(n_blocks)]
        for _ in range(2):  # upsampling
            layers += [nn.ConvTranspose2d(ch, ch // 2, kernel_size=3, stride=2, padding=1, output_padding=1,
                                          bias=use_bias),
                       norm_layer(ch // 2), nn.ReLU(True)]
            ch //= 2
        layers += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*layers)

    def _body(self):
        return self.model


class ResnetBlock(nn.Module):
    """models/networks.py:195-236: x + conv_block(x) with conv_block = pad-conv-norm-relu-[dropout]-pad-conv-norm."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super(ResnetBlock, self).__init__()
        self.conv_block = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias)

    def build_conv_block(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        def padded_conv():
            if padding_type == 'reflect':
                return [nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias)]
            if padding_type == 'replicate':
                return [nn.ReplicationPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias)]
            if padding_type == 'zero':
                return [nn.Conv2d(dim, dim, kernel_size=3, padding=1, bias=use_bias)]
            raise NotImplementedError('padding [%s] is not implemented' % padding_type)

        block = padded_conv() + [norm_layer(dim), nn.ReLU(True)]
        if use_dropout:
            block += [nn.Dropout(0.5)]
        block += padded_conv() + [norm_layer(dim)]
        return nn.Sequential(*block)

    def _block_body(self, tape, x):
        stages, _ = engine.compile_chain([self])
        a, b = stages
        v0 = tape.input_nchw(x, pad=a.reflect, pad_kind='reflect' if a.reflect else None,
                             want_grad=tape.input_wants[0])
        if not a.reflect and a.conv.padding[0]:
            raise NotImplementedError("standalone ResnetBlock with padding_type='zero'")
        mid = tape.stage(v0, a.conv, a.norm, a.act, a.slope, reflect=a.reflect, halo=b.reflect,
                         halo_kind='reflect' if b.reflect else None)
        y = tape.stage(mid, b.conv, b.norm, b.act, b.slope, res=v0, reflect=b.reflect)
        out, slot = tape.output_nchw(y)
        return [out], [slot], [v0]

    def forward(self, x):
        """models/networks.py:234-236: out = x + conv_block(x).  Inside a ResnetGenerator the block is a pair of fused
        stages of the generator's plan; called on its own (fp32 NCHW in / out) it runs the same two stages."""
        return graph.run(self, self._block_body, [x])[0]


class NLayerDiscriminator(_FusedNet):
    """models/networks.py:320-364: the 70x70 PatchGAN."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False):
        super(NLayerDiscriminator, self).__init__()
        use_bias = _bias_follows_instance_norm(norm_layer)
        kw, padw = 4, 1
        seq = [nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers + 1):
            prev, mult = mult, min(2 ** n, 8)
            stride = 2 if n < n_layers else 1
            seq += [nn.Conv2d(ndf * prev, ndf * mult, kernel_size=kw, stride=stride, padding=padw, bias=use_bias),
                    norm_layer(ndf * mult), nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(ndf * mult, 1, kernel_size=kw, stride=1, padding=padw)]
        if use_sigmoid:
            seq += [nn.Sigmoid()]
        self.model = nn.Sequential(*seq)

    def _body(self):
        return self.model


class PixelDiscriminator(_FusedNet):
    """models/networks.py:367-389: three 1x1 convolutions."""

    def __init__(self, input_nc, ndf=64, norm_layer=nn.BatchNorm2d, use_sigmoid=False):
        super(PixelDiscriminator, self).__init__()
        use_bias = _bias_follows_instance_norm(norm_layer)
        net = [nn.Conv2d(input_nc, ndf, kernel_size=1, stride=1, padding=0), nn.LeakyReLU(0.2, True),
               nn.Conv2d(ndf, ndf * 2, kernel_size=1, stride=1, padding=0, bias=use_bias), norm_layer(ndf * 2),
               nn.LeakyReLU(0.2, True), nn.Conv2d(ndf * 2, 1, kernel_size=1, stride=1, padding=0, bias=use_bias)]
        if use_sigmoid:
            net.append(nn.Sigmoid())
        self.net = nn.Sequential(*net)

    def _body(self):
        return self.net


class UnetGenerator(nn.Module):
    """models/networks.py:243-260. Module tree / state_dict identical to the reference; ``forward`` runs the
    whole U-Net on the graph engine (``graph.py``):

    * every skip concatenation ``cat([x, model(x)], 1)`` (:316) is one preallocated NHWC buffer whose two
      channel halves are written by their producers — ``torch.cat`` never runs;
    * the skip half carries LeakyReLU(x) because the child's in-place ``downrelu`` mutates x before the
      concatenation materialises (:279,:316, SURVEY B-5), and the parent's in-place ``uprelu`` then acts on the
      whole concatenation: the buffer therefore stores relu(leaky(x)) = relu(x) next to relu(up-path);
    * down path: conv -> norm -> LeakyReLU fused per level; up path: convT -> norm -> ReLU (-> dropout)."""

    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False):
        super(UnetGenerator, self).__init__()
        block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=None, norm_layer=norm_layer,
                                        innermost=True)
        for _ in range(num_downs - 5):
            block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=block,
                                            norm_layer=norm_layer, use_dropout=use_dropout)
        for outer, inner in ((ngf * 4, ngf * 8), (ngf * 2, ngf * 4), (ngf, ngf * 2)):
            block = UnetSkipConnectionBlock(outer, inner, input_nc=None, submodule=block, norm_layer=norm_layer)
        self.model = UnetSkipConnectionBlock(output_nc, ngf, input_nc=input_nc, submodule=block, outermost=True,
                                             norm_layer=norm_layer)

    def _levels(self):
        """[(downconv, downnorm, upconv, upnorm, dropout)] from the outermost block inwards."""
        levels, b = [], self.model
        while b is not None:
            mods = list(b.model.children())
            sub = [m for m in mods if isinstance(m, UnetSkipConnectionBlock)]
            down = [m for m in mods if isinstance(m, nn.Conv2d)][0]
            up = [m for m in mods if isinstance(m, nn.ConvTranspose2d)][0]
            norms = [m for m in mods if isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d))]
            drop = [m for m in mods if isinstance(m, nn.Dropout)]
            downnorm = upnorm = None
            if b.outermost:
                pass
            elif b.innermost:
                upnorm = norms[0]
            else:
                downnorm, upnorm = norms[0], norms[1]
            levels.append((down, downnorm, up, upnorm, drop[0] if drop else None))
            b = sub[0] if sub else None
        return levels

    def _body(self, tape, x):
        lv = self._levels()
        L = len(lv)
        n = x.shape[0]
        img = tape.input_nchw(x, first_conv=lv[0][0], want_grad=tape.input_wants[0])
        # ---- down path: d[j] = LeakyReLU(norm_j(conv_j(d[j-1]))); skips[j] = concat buffer read by upconv_{j-1}
        cur = img
        skips = [None] * L
        for j in range(L - 1):
            down, downnorm = lv[j][0], lv[j][1]
            cur = tape.stage(cur, down, downnorm, ACT_LEAKY, 0.2)
            _, h, w, _ = cur.t.shape
            c = cur.c
            cat = tape.concat_buffer(n, h, w, 2 * c)
            tape.norm_act(cur, None, ACT_RELU, out=cat.slice(0, c))     # relu(leaky(x)) half of the skip
            skips[j + 1] = cat
        # ---- innermost: relu(downconv(.)) -> upconv -> upnorm
        down, _, up, upnorm, _ = lv[L - 1]
        inner = tape.stage(cur, down, None, ACT_RELU)
        c = skips[L - 1].c // 2
        tape.stage(inner, up, upnorm, ACT_RELU, out=skips[L - 1].slice(c, 2 * c))
        # ---- up path
        for j in range(L - 2, 0, -1):
            _, _, up, upnorm, drop = lv[j]
            c = skips[j].c // 2
            dst = skips[j].slice(c, 2 * c)
            tape.stage(skips[j + 1], up, upnorm, ACT_RELU, out=dst)
            if drop is not None:
                tape.dropout(dst, float(drop.p))
        out, slot = tape.stage(skips[1], lv[0][2], None, ACT_TANH, out_nchw=True)
        return [out], [slot], [img]

    def forward(self, input):
        return graph.run(self, self._body, [input])[0]


class UnetSkipConnectionBlock(nn.Module):
    """models/networks.py:266-316 (parameter layout only, see UnetGenerator)."""

    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=nn.BatchNorm2d, use_dropout=False):
        super(UnetSkipConnectionBlock, self).__init__()
        self.outermost = outermost
        self.innermost = innermost
        use_bias = _bias_follows_instance_norm(norm_layer)
        if input_nc is None:
            input_nc = outer_nc
        downconv = nn.Conv2d(input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        downrelu, uprelu = nn.LeakyReLU(0.2, True), nn.ReLU(True)
        downnorm, upnorm = norm_layer(inner_nc), norm_layer(outer_nc)
        if outermost:
            upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            model = [downconv, submodule, uprelu, upconv, nn.Tanh()]
        elif innermost:
            upconv = nn.ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            model = [downrelu, downconv, uprelu, upconv, upnorm]
        else:
            upconv = nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            model = [downrelu, downconv, downnorm, submodule, uprelu, upconv, upnorm]
            if use_dropout:
                model += [nn.Dropout(0.5)]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        raise RuntimeError("UnetSkipConnectionBlock runs as part of a fused cdb200 network (UnetGenerator.forward)")
