"""General NHWC graph engine: a define-by-run tape of fused B200 ops with a hand-written backward.

``engine.py`` executes straight conv chains (ResnetGenerator, PatchGAN).  The U-Net of pix2pix
(models/networks.py:243-316) and the seg/depth networks (new_multi/networks5_ds.py) have skip
connections, channel concatenations, pre-activation BatchNorm, pooling, attention gates and
re-sampling; they run on this tape instead.  A network's ``forward`` is written once against the tape
API below; every call records closures that the single ``torch.autograd.Function`` of the network
call replays in reverse.

Values (``Val``) are NHWC bf16 views, possibly channel slices of a shared concatenation buffer and
possibly surrounded by a materialised zero / reflect halo.  Gradients are lists of bf16 contributions
summed by the consuming kernel (``norm_act_bwd`` takes two) or, for dense-block buffers with dozens of
consumers, one fp32 accumulator that the BatchNorm backward kernels add into directly.
"""
import torch
import torch.nn as nn

from . import engine, ops
from .ops import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, NORM_BATCH, NORM_FLAG_ACCUM_F32,
                  NORM_FLAG_ACT_FIRST, NORM_INSTANCE, NORM_NONE)

BF16 = torch.bfloat16
_pack = engine._pack_cache


class Val:
    """An activation: ``t`` = interior view [N,H,W,Cstore] (unit channel stride), ``c`` real channels."""

    def __init__(self, t, c, buf=None, halo=0, halo_kind=None):
        self.t = t
        self.c = c
        self.buf = buf if buf is not None else t
        self.halo = halo
        self.halo_kind = halo_kind       # 'zero' | 'reflect' | None
        self.grads = []                  # [(bf16 view shaped like t, covers_halo: bool)]
        self.grad32 = None               # fp32 accumulator view shaped like t (dense-block buffers)
        self.stats = None                # [1, C, 2] batch sums maintained by the producers (dense blocks)
        self.parent = None               # concatenation buffer this value is a channel slice of
        self.c0 = 0
        self.rowpack = 0
        self.folded_parent = False   # the parent's contributions were already merged into `grads`
        self.is_input = False
        self.want_grad = False

    @property
    def shape(self):
        return self.t.shape

    def slice(self, c0, c1):
        """Channel slice [c0, c1) (multiples of 8) sharing storage, gradient accumulator and statistics."""
        assert c0 % 8 == 0 and (c1 % 8 == 0 or c1 == self.c)
        v = Val(self.t[..., c0:ops.round_up(c1, 8)], c1 - c0)
        v.parent, v.c0 = self, c0
        if self.grad32 is not None:
            v.grad32 = self.grad32[..., c0:ops.round_up(c1, 8)]
        if self.stats is not None:
            v.stats = self.stats[:, c0:c1]
        return v


def _act_of(m):
    return engine._act_of(m)


def _regular_pitch(t):
    """[N,H,W,C] view whose pixels are evenly pitched (contiguous, or a channel prefix of a wider buffer): the
    flat convolution kernel addresses it as a 2-D matrix [N*H*W, C] with row pitch stride(2)."""
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    return sc == 1 and sw >= c and sh == w * sw and sn == h * w * sw


class Tape:
    def __init__(self, training, device, record):
        self.training = training
        self.dev = device
        self.record = record
        self.back = []
        self.param_grads = {}
        self.needs = {}            # param -> bool, filled in before the backward replay
        self.arena = ops.ZeroArena(device)
        # storage / arithmetic mode of this network call (ops.set_precision): bf16 storage + kind::f16, or fp32
        # storage + kind::tf32 (single pass, or error-compensated 'tf32x3')
        self.prec = ops.get_precision()
        self.dtype = BF16 if self.prec == 'bf16' else torch.float32
        # data parallelism: BatchNorm layers normalise over the shards of all ranks (statistics all-reduced per layer)
        self.bn_world = ops.bn_world()

    # ---------------------------------------------------------------- convolution calls (precision dispatch)
    def _operand(self, x, mode_x3, mode_tf32=3):
        return ops.split_tf32(x, mode_x3 if self.prec == 'tf32x3' else mode_tf32)

    def _conv(self, g, xin, weight, rows_are_dim0, out, bias=None, act=ACT_NONE, slope=0.0, stats=None,
              stats_batch=False, rowpack=0, flipped=False):
        """One forward / data-gradient convolution launch in the precision of this call."""
        if self.prec == 'bf16':
            wp, rows_pad, kpad = _pack.get(weight, rows_are_dim0, rowpack, flipped)
            ops.conv2d_fwd_ex(g, xin, wp, rows_pad, kpad, out, bias, act, slope, stats, stats_batch)
            return
        assert not rowpack
        wp, rows_pad, kpad = _pack.get_tf32(weight, rows_are_dim0, xin.shape[3], self.prec == 'tf32x3', flipped)
        ops.conv2d_fwd_ex(g, self._operand(xin, 0), wp, rows_pad, kpad, out, bias, act, slope, stats, stats_batch)

    def _wgrad(self, g, x, dy, dw):
        if self.prec == 'bf16':
            ops.conv2d_wgrad(g, x, dy, dw, False)
        else:
            ops.conv2d_wgrad(g, self._operand(x, 1), self._operand(dy, 2), dw, False)

    # ---------------------------------------------------------------- bookkeeping
    def add_param_grad(self, p, g):
        if p is None or not self.needs.get(p, False):
            return
        cur = self.param_grads.get(p)
        self.param_grads[p] = g if cur is None else cur + g

    def wants(self, p):
        return p is not None and self.needs.get(p, False)

    def add_grad(self, v, g, covers_halo=False):
        """Registers a gradient contribution for v (g shaped like v.t; with covers_halo it is the interior
        view of a gradient of the whole padded buffer)."""
        if v.grad32 is not None:
            ops.cast(g, v.grad32, accumulate=True)
            return
        v.grads.append((g, covers_halo))

    def gather(self, v):
        """Gradient contributions of v as (dout, dskip): at most one halo-covering contribution goes to
        dout; extra contributions are summed with the add kernel."""
        items = list(v.grads)
        if v.parent is not None and v.grad32 is None and not v.folded_parent:
            for g, ch in v.parent.grads:
                items.append((g[..., v.c0:v.c0 + v.t.shape[3]], False))
        if v.grad32 is not None:
            g = torch.empty(tuple(v.t.shape), dtype=self.dtype, device=self.dev)
            ops.cast(v.grad32, g)
            items.append((g, False))
        if not items:
            return None, None
        halo_items = [g for g, ch in items if ch]
        plain = [g for g, ch in items if not ch]
        while len(halo_items) > 1:
            # several consumers read the reflect-padded buffer (models/encoder_decoder.py:196-197: a decoder level and
            # its output block): fold the halo of the extra contributions into interior-shaped tensors first
            if v.halo_kind != 'reflect':
                raise NotImplementedError("two halo-covering gradient contributions of a zero-padded value")
            g = halo_items.pop()
            folded = torch.empty(tuple(v.t.shape), dtype=self.dtype, device=self.dev)
            desc = ops.norm_desc(NORM_NONE, ACT_NONE, 0.0, 0.0, v.c, v.halo)
            ops.norm_act_bwd(desc, v.t, folded, g, None, None, None)
            plain.append(folded)
        dout = halo_items[0] if halo_items else None
        while len(plain) > (1 if dout is not None else 2):
            a, b = plain.pop(), plain.pop()
            s = torch.empty(tuple(a.shape), dtype=self.dtype, device=self.dev)
            ops.add(a, b, s)
            plain.append(s)
        if dout is None:
            dout = plain.pop(0)
        dskip = plain[0] if plain else None
        return dout, dskip

    def total_grad(self, v):
        """One tensor holding the summed gradient of v (interior), or None."""
        dout, dskip = self.gather(v)
        if dout is None:
            return None
        if dskip is None and not (v.halo and v.halo_kind == 'reflect' and any(ch for _, ch in v.grads)):
            return dout
        out = torch.empty(tuple(v.t.shape), dtype=self.dtype, device=self.dev)
        desc = ops.norm_desc(NORM_NONE, ACT_NONE, 0.0, 0.0, v.c, v.halo if v.halo_kind == 'reflect' else 0)
        ops.norm_act_bwd(desc, v.t, out, dout, dskip, None, None)
        return out

    # ---------------------------------------------------------------- allocation
    def new_val(self, n, h, w, c, halo=0, halo_kind=None, slack_w=0):
        cs = ops.round_up(c, 8)
        if halo and halo_kind == 'zero':
            buf = ops.empty_zero_halo(n, h, w, cs, halo, slack_w, self.dev, self.dtype)
        else:
            buf = torch.empty((n, h + 2 * halo, w + 2 * halo + slack_w, cs), dtype=self.dtype, device=self.dev)
        t = buf[:, halo:halo + h, halo:halo + w, :]
        full = buf if slack_w == 0 else buf[:, :, :w + 2 * halo, :]
        return Val(t, c, full, halo, halo_kind if halo else None)

    def concat_buffer(self, n, h, w, c, f32grad=False, stats=False, halo=0, halo_kind=None):
        """Preallocated concatenation buffer whose channel slices are written by the producers."""
        v = self.new_val(n, h, w, c, halo, halo_kind)
        if f32grad and self.record:
            v.grad32 = torch.zeros((n, h, w, ops.round_up(c, 8)), dtype=torch.float32, device=self.dev)
        if stats:
            v.stats = self.arena.take((1, c, 2))
        return v

    # ---------------------------------------------------------------- module boundary
    def input_nchw(self, x, pad=0, pad_kind=None, first_conv=None, want_grad=False, out=None):
        """fp32 NCHW tensor -> Val (with the first convolution's padding materialised). Returns the Val;
        the gradient w.r.t. x (fp32 NCHW) is produced by input_grad() during the backward replay."""
        if not x.is_cuda:
            raise RuntimeError("cdb200 networks run on CUDA tensors only (no CPU path)")
        if x.dtype != torch.float32 or x.dim() != 4:
            raise TypeError("fp32 NCHW input expected")
        n, c, h, w = x.shape
        rp = 0
        if self.prec == 'bf16' and first_conv is not None and not isinstance(first_conv, nn.ConvTranspose2d):
            k = first_conv.kernel_size[0]
            if first_conv.dilation[0] == 1 and ((c <= 8 and k <= 8) or (c <= 16 and k <= 4)):
                rp = 8 if c <= 8 else 16
        if rp:
            conv = first_conv
            p0 = pad if pad else conv.padding[0]
            kind = pad_kind if pad else 'zero'
            k, st = conv.kernel_size[0], conv.stride[0]
            wo = (w + 2 * p0 - (k - 1) - 1) // st + 1
            wneed = max(w + 2 * p0, st * (wo - 1) + 64 // rp)
            buf = torch.zeros((n, h + 2 * p0, wneed, rp), dtype=self.dtype, device=self.dev)
            t = buf[:, p0:p0 + h, p0:p0 + w, :]
            ops.nchw_to_nhwc(x, t, pad=p0 if kind == 'reflect' else 0)
            v = Val(t, c, buf, p0, kind if p0 else None)
            v.rowpack = rp
        elif out is not None:
            v = out
            ops.nchw_to_nhwc(x, v.t, pad=0)
            self._fill_stats(v)
        else:
            v = self.new_val(n, h, w, c, pad, pad_kind)
            ops.nchw_to_nhwc(x, v.t, pad=pad if pad_kind == 'reflect' else 0)
        v.is_input = True
        v.want_grad = want_grad
        return v

    def _fill_stats(self, v):
        if v.stats is not None:
            ops.channel_stats(v.t, v.c, False, v.stats)

    def input_grad(self, v, shape):
        """fp32 NCHW gradient of a network input (after the backward replay)."""
        g = self.total_grad(v)
        if g is None:
            return None
        out = torch.empty(shape, dtype=torch.float32, device=self.dev)
        ops.nhwc_to_nchw(g, shape[1], out)
        return out

    def output_nchw(self, v, differentiable=True):
        """Val -> fp32 NCHW tensor. Registers the slot that receives the incoming gradient."""
        n, h, w, _ = v.t.shape
        out = torch.empty((n, v.c, h, w), dtype=torch.float32, device=self.dev)
        ops.nhwc_to_nchw(v.t, v.c, out)
        slot = {'val': v, 'kind': 'val', 'differentiable': differentiable}
        return out, slot

    def seed_output_grad(self, slot, gout):
        if gout is None or not slot['differentiable']:
            return
        gout = gout.contiguous()
        if slot['kind'] == 'val':
            v = slot['val']
            g = torch.empty(tuple(v.t.shape), dtype=self.dtype, device=self.dev)
            ops.nchw_to_nhwc(gout, g, pad=0)
            self.add_grad(v, g)
        else:
            slot['gout'] = gout

    # ---------------------------------------------------------------- convolution stage
    def _conv_operand(self, x, conv, transposed, reflect):
        """(input tensor handed to the kernel, effective padding, materialised?)."""
        p = conv.padding[0]
        if reflect:
            if not (x.halo == reflect and x.halo_kind == 'reflect' and p == 0 and not transposed):
                raise NotImplementedError("reflect padding must be materialised by the producer")
            return x.buf, 0, True
        if x.rowpack:
            return x.buf, 0, True
        if (not transposed and p > 0 and conv.stride[0] == 1 and x.halo == p and x.halo_kind == 'zero'):
            return x.buf, 0, True
        return x.t, p, False

    @staticmethod
    def _out_hw(conv, transposed, h, w, reflect):
        k, s, p, d = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
        if transposed:
            op = conv.output_padding[0]
            return (h - 1) * s - 2 * p + d * (k - 1) + op + 1, (w - 1) * s - 2 * p + d * (k - 1) + op + 1
        hp, wp = h + 2 * reflect, w + 2 * reflect
        return (hp + 2 * p - d * (k - 1) - 1) // s + 1, (wp + 2 * p - d * (k - 1) - 1) // s + 1

    @staticmethod
    def _check_conv(conv):
        if conv.groups != 1:
            raise NotImplementedError("grouped convolution")
        if (conv.kernel_size[0] != conv.kernel_size[1] or conv.stride[0] != conv.stride[1]
                or conv.padding[0] != conv.padding[1] or conv.dilation[0] != conv.dilation[1]):
            raise NotImplementedError("non-square convolution geometry")
        if conv.stride[0] not in (1, 2):
            raise NotImplementedError("stride %d" % conv.stride[0])
        if getattr(conv, "padding_mode", "zeros") != "zeros":
            raise NotImplementedError("padding_mode %s" % conv.padding_mode)

    def stage(self, x, conv, norm=None, act=ACT_NONE, slope=0.0, act_first=False, res=None, reflect=0,
              out=None, halo=0, halo_kind=None, out_nchw=False, differentiable_out=True):
        """[reflect pad] -> conv -> [act] -> [norm] -> [act] -> [+ res], fused as in engine.py.

        act_first: conv -> act -> norm (the R_dep blocks of networks5_ds.py).  out: an existing Val (channel
        slice of a concatenation buffer) to write into.  out_nchw: the stage is a network output written as
        fp32 NCHW by the convolution epilogue (no norm); returns (tensor, slot)."""
        self._check_conv(conv)
        transposed = isinstance(conv, nn.ConvTranspose2d)
        n, hi, wi, _ = x.t.shape
        ci = x.c
        k, stride, dil = conv.kernel_size[0], conv.stride[0], conv.dilation[0]
        co = conv.out_channels
        cs = ops.round_up(co, 8)
        ho, wo = self._out_hw(conv, transposed, hi, wi, reflect)
        xin, pad, materialised = self._conv_operand(x, conv, transposed, reflect)
        rowpack = x.rowpack
        g = ops.geom(k, k, stride, pad, pad, dil, transposed, rowpack)
        flat = (not transposed and stride == 1 and not rowpack and pad == 0 and _regular_pitch(xin))
        if norm is None:
            nk = NORM_NONE
        elif isinstance(norm, nn.InstanceNorm2d):
            if norm.affine or norm.track_running_stats:
                raise NotImplementedError("InstanceNorm2d with affine / running stats")
            nk = NORM_INSTANCE
        else:
            nk = NORM_BATCH
        training = self.training
        use_running = nk == NORM_BATCH and not training and norm.track_running_stats

        if out_nchw:
            assert nk == NORM_NONE and res is None
            o = torch.empty((n, co, ho, wo), dtype=torch.float32, device=self.dev)
            self._conv(g, xin, conv.weight, not transposed, ops.out_view_nchw(o), conv.bias, act, slope, rowpack=rowpack)
            slot = {'kind': 'nchw', 'differentiable': differentiable_out, 'gout': None}
            outv = None
        else:
            outv = out if out is not None else self.new_val(n, ho, wo, co, halo, halo_kind)
            assert tuple(outv.t.shape[:3]) == (n, ho, wo) and outv.c == co
            slot = None

        y = stats = None
        if out_nchw:
            pass
        elif nk == NORM_NONE:
            if res is not None:
                raise NotImplementedError("residual add without normalisation")
            if outv.halo and outv.halo_kind == 'reflect':
                raise NotImplementedError("reflect halo after a stage without normalisation")
            self._conv(g, xin, conv.weight, not transposed, ops.out_view_nhwc(outv.t, co), conv.bias, act, slope,
                       outv.stats, stats_batch=True, rowpack=rowpack)
        else:
            if flat:
                y = ops.alloc_flat_output(n, ho, wo, xin.shape[2], cs, self.dev, dtype=self.dtype)
            else:
                y = torch.empty((n, ho, wo, cs), dtype=self.dtype, device=self.dev)
            # a bias directly in front of a batch-statistics normalisation cancels; with an activation in
            # between (act_first) or running statistics it does not
            bias = conv.bias if (use_running or act_first) else None
            if not use_running:
                groups = n if nk == NORM_INSTANCE else 1
                stats = self.arena.take((groups, co, 2))
            self._conv(g, xin, conv.weight, not transposed, ops.out_view_nhwc(y, co), bias,
                       act if act_first else ACT_NONE, slope, stats, stats_batch=(nk == NORM_BATCH), rowpack=rowpack)
            if nk == NORM_BATCH and stats is not None and self.bn_world > 1:
                ops.bn_all_reduce(stats)          # batch statistics over the shards of all ranks
            affine = getattr(norm, "affine", False)
            desc = self._norm_desc(nk, norm, act, slope, co, outv.halo if outv.halo_kind == 'reflect' else 0, stats,
                                   use_running, update=True, flags=NORM_FLAG_ACT_FIRST if act_first else 0,
                                   conv_bias=conv.bias if bias is None else None)
            ops.norm_act_fwd(desc, y, outv.t, res.t if res is not None else None)
            if outv.stats is not None:
                ops.channel_stats(outv.t, co, False, outv.stats)

        if self.record:
            def backward():
                self._stage_backward(x, conv, transposed, norm, nk, act, slope, act_first, res, reflect, outv, slot, o
                                     if out_nchw else None, y, stats, xin, pad, materialised, rowpack, flat, use_running,
                                     (n, hi, wi, ci, ho, wo, co, cs))
            self.back.append(backward)
        if out_nchw:
            return o, slot
        return outv

    def _norm_desc(self, nk, norm, act, slope, co, pad, stats, use_running, update=False, flags=0, conv_bias=None):
        affine = norm is not None and getattr(norm, "affine", False)
        bn = nk == NORM_BATCH
        upd = bool(update and bn and self.training and norm.track_running_stats)
        if upd and norm.num_batches_tracked is not None:
            norm.num_batches_tracked += 1
        mom = 0.1
        if norm is not None and getattr(norm, "momentum", None) is not None:
            mom = float(norm.momentum)
        return ops.norm_desc(nk, act, slope, float(norm.eps) if norm is not None else 0.0, co, pad, stats,
                             norm.weight if affine else None, norm.bias if affine else None,
                             norm.running_mean if bn else None, norm.running_var if bn else None,
                             use_running=use_running, update_running=upd, momentum=mom, flags=flags,
                             conv_bias=conv_bias, count_scale=self.bn_world if (bn and not use_running) else 1)

    def _stage_backward(self, x, conv, transposed, norm, nk, act, slope, act_first, res, reflect, outv, slot, o_nchw,
                        y, stats, xin, pad, materialised, rowpack, flat, use_running, dims):
        n, hi, wi, ci, ho, wo, co, cs = dims
        dev = self.dev
        k, dil, stride = conv.kernel_size[0], conv.dilation[0], conv.stride[0]
        want_w = self.wants(conv.weight)
        want_b = self.wants(conv.bias)
        want_dx = (not x.is_input) or x.want_grad
        affine = norm is not None and getattr(norm, "affine", False)
        flat_dgrad = flat and want_dx and materialised
        # ---- gradient w.r.t. the raw convolution output
        if flat_dgrad:
            hz = (k - 1) * dil
            slack = 64 // cs if cs <= 16 else 0
            dyp = ops.empty_zero_halo(n, ho, wo, cs, hz, slack, dev, self.dtype)
            dy = dyp[:, hz:hz + ho, hz:hz + wo, :]
        else:
            dyp = None
            dy = None
        if slot is not None:                       # fp32 NCHW network output
            gout = slot.get('gout')
            if gout is None:
                return
            if dy is None:
                dy = torch.empty((n, ho, wo, cs), dtype=self.dtype, device=dev)
            ops.nchw_to_nhwc(gout, dy, pad=0, act_out=o_nchw if act != ACT_NONE else None, act=act, slope=slope)
            if want_b:
                db = torch.empty((co,), dtype=torch.float32, device=dev)
                ops.bias_grad_nchw(gout, o_nchw if act != ACT_NONE else None, act, slope, db)
                self.add_param_grad(conv.bias, db)
        else:
            dout, dskip = self.gather(outv)
            if dout is None:
                return
            halo_fold = outv.halo if (outv.halo_kind == 'reflect') else 0
            if nk == NORM_NONE:
                need_kernel = (act != ACT_NONE) or want_b or dskip is not None or halo_fold or flat_dgrad
                if need_kernel:
                    if dy is None:
                        dy = torch.empty((n, ho, wo, cs), dtype=self.dtype, device=dev)
                    bstats = self.arena.take((1, co, 2)) if want_b else None
                    desc = ops.norm_desc(NORM_NONE, act, slope, 0.0, co, halo_fold)
                    if act in (ACT_TANH, ACT_SIGMOID):
                        raise NotImplementedError("tanh / sigmoid inside a network (only at fp32 outputs)")
                    ops.norm_act_bwd(desc, outv.t, dy, dout, dskip, bstats, None)
                    if want_b:
                        self.add_param_grad(conv.bias, bstats[0, :, 0].contiguous())
                else:
                    dy = dout
            else:
                if dy is None:
                    dy = torch.empty((n, ho, wo, cs), dtype=self.dtype, device=dev)
                groups = n if nk == NORM_INSTANCE else 1
                bstats = self.arena.take((groups, co, 2))
                base_flags = NORM_FLAG_ACT_FIRST if act_first else 0
                gsum = None
                if res is not None and (dskip is not None or halo_fold):
                    gsum = torch.empty((n, ho, wo, cs), dtype=self.dtype, device=dev)
                if nk == NORM_BATCH and not use_running and self.bn_world > 1:
                    ops.norm_act_bwd_synced(
                        lambda extra: self._norm_desc(nk, norm, act, slope, co, halo_fold, stats, use_running,
                                                      flags=base_flags | extra),
                        y, dy, dout, dskip, bstats, gsum, self.bn_world)
                else:
                    desc = self._norm_desc(nk, norm, act, slope, co, halo_fold, stats, use_running, flags=base_flags)
                    ops.norm_act_bwd(desc, y, dy, dout, dskip, bstats, gsum)
                if res is not None:
                    self.add_grad(res, gsum if gsum is not None else dout)
                if want_b:
                    if use_running or act_first:
                        bs = self.arena.take((1, co, 2))
                        ops.channel_stats(dy, co, False, bs)
                        self.add_param_grad(conv.bias, bs[0, :, 0].contiguous())
                    else:
                        self.add_param_grad(conv.bias, torch.zeros_like(conv.bias))
                if affine:
                    self.add_param_grad(norm.weight, bstats[0, :, 1].contiguous())
                    self.add_param_grad(norm.bias, bstats[0, :, 0].contiguous())
        # ---- weight gradient
        if want_w:
            if self.prec == 'bf16' and flat_dgrad and cs <= 16 and k * cs <= 64 and dil == 1:
                tmp = torch.empty((ci, co, k, k), dtype=torch.float32, device=dev)
                ops.conv2d_wgrad(ops.geom(k, k, 1, 0, 0, 1, True, cs), xin, dyp, tmp, False)
                dw = tmp.flip(2, 3).permute(1, 0, 2, 3).contiguous()
            else:
                dw = torch.empty_like(conv.weight, memory_format=torch.contiguous_format)
                self._wgrad(ops.geom(k, k, stride, pad, pad, dil, transposed, rowpack), xin, dy, dw)
            self.add_param_grad(conv.weight, dw)
        # ---- data gradient
        if not want_dx:
            return
        if rowpack:
            # the row-packed image buffer is zero padded by construction: gradient of the interior only
            p = conv.padding[0] if not reflect else 0
            gd = ops.geom(k, k, stride, p, p, dil, not transposed, 0)
            if reflect:
                full = torch.empty((n, hi + 2 * reflect, wi + 2 * reflect, ops.round_up(ci, 8)), dtype=self.dtype, device=dev)
                self._conv(gd, dy, conv.weight, transposed, ops.out_view_nhwc(full, ci))
                self.add_grad(x, full[:, reflect:reflect + hi, reflect:reflect + wi, :], covers_halo=True)
            else:
                dx = torch.empty((n, hi, wi, ops.round_up(ci, 8)), dtype=self.dtype, device=dev)
                self._conv(gd, dy, conv.weight, transposed, ops.out_view_nhwc(dx, ci))
                self.add_grad(x, dx)
            return
        if flat_dgrad:
            hp, wp_ = xin.shape[1], xin.shape[2]
            dfull = ops.alloc_flat_output(n, hp, wp_, dyp.shape[2], xin.shape[3], dev, dtype=self.dtype)
            if self.prec == 'bf16' and cs <= 16 and k * cs <= 64 and dil == 1:
                # few output channels: row-packed dy operand (one K block per filter row), taps reversed at packing
                self._conv(ops.geom(k, k, 1, 0, 0, 1, False, cs), dyp, conv.weight, transposed,
                           ops.out_view_nhwc(dfull, ci), rowpack=cs, flipped=True)
            else:
                gflip = ops.geom(k, k, 1, 0, 0, dil, False, 0, True)
                self._conv(gflip, dyp, conv.weight, transposed, ops.out_view_nhwc(dfull, ci))
        elif k == 1 and stride == 1 and pad == 0 and not transposed and _regular_pitch(dy):
            # the data gradient of a 1x1 convolution is a 1x1 convolution: flat kernel (plain GEMM, BM = 256)
            dfull = ops.alloc_flat_output(n, hi, wi, wi, ops.round_up(ci, 8), dev, dtype=self.dtype)
            self._conv(ops.geom(1, 1), dy, conv.weight, transposed, ops.out_view_nhwc(dfull, ci))
        else:
            gd = ops.geom(k, k, stride, pad, pad, dil, not transposed, 0)
            dfull = torch.empty(tuple(xin.shape), dtype=self.dtype, device=dev)
            self._conv(gd, dy, conv.weight, transposed, ops.out_view_nhwc(dfull, ci))
        if materialised:
            h_ = x.halo
            self.add_grad(x, dfull[:, h_:h_ + hi, h_:h_ + wi, :], covers_halo=(x.halo_kind == 'reflect'))
        else:
            self.add_grad(x, dfull)

    # ---------------------------------------------------------------- standalone norm / activation
    def norm_act(self, x, norm, act=ACT_NONE, slope=0.0, halo=0, halo_kind=None, out=None):
        """out = act(norm(x)) for a value that already exists (pre-activation BatchNorm of the dense layers,
        networks5_ds.py:125-131; norm may be None for a pure activation / re-padding copy)."""
        n, h, w, _ = x.t.shape
        c = x.c
        if norm is None:
            nk = NORM_NONE
        elif isinstance(norm, nn.InstanceNorm2d):
            nk = NORM_INSTANCE
        else:
            nk = NORM_BATCH
        use_running = nk == NORM_BATCH and not self.training and norm.track_running_stats
        stats = None
        if nk != NORM_NONE and not use_running:
            if nk == NORM_BATCH and x.stats is not None:
                stats = x.stats
                if self.bn_world > 1:
                    stats = x.stats.clone()       # the producers' slice holds this rank's sums; consumers sum a copy
                    ops.bn_all_reduce(stats)
            else:
                groups = n if nk == NORM_INSTANCE else 1
                stats = self.arena.take((groups, c, 2))
                ops.channel_stats(x.t, c, nk == NORM_INSTANCE, stats)
                if nk == NORM_BATCH and self.bn_world > 1:
                    ops.bn_all_reduce(stats)
        outv = out if out is not None else self.new_val(n, h, w, c, halo, halo_kind)
        desc = self._norm_desc(nk, norm, act, slope, c, outv.halo if outv.halo_kind == 'reflect' else 0, stats,
                               use_running, update=True)
        ops.norm_act_fwd(desc, x.t, outv.t, None)
        if outv.stats is not None:
            ops.channel_stats(outv.t, c, False, outv.stats)
        if self.record:
            def backward():
                dout, dskip = self.gather(outv)
                if dout is None:
                    return
                if act in (ACT_TANH, ACT_SIGMOID):
                    raise NotImplementedError("tanh / sigmoid inside a network")
                affine = norm is not None and getattr(norm, "affine", False)
                groups = n if nk == NORM_INSTANCE else 1
                bstats = None
                if nk != NORM_NONE and not use_running or affine:
                    bstats = self.arena.take((groups, c, 2))
                fold = outv.halo if outv.halo_kind == 'reflect' else 0
                flags = NORM_FLAG_ACCUM_F32 if x.grad32 is not None else 0
                synced = nk == NORM_BATCH and not use_running and self.bn_world > 1

                def run_bwd(dy_target):
                    if synced:
                        ops.norm_act_bwd_synced(
                            lambda extra: self._norm_desc(nk, norm, act, slope, c, fold, stats, use_running,
                                                          flags=flags | extra),
                            x.t, dy_target, dout, dskip, bstats, None, self.bn_world)
                    else:
                        desc_b = self._norm_desc(nk, norm, act, slope, c, fold, stats, use_running, flags=flags)
                        ops.norm_act_bwd(desc_b, x.t, dy_target, dout, dskip, bstats, None)

                if x.grad32 is not None:
                    run_bwd(x.grad32)
                else:
                    dy = torch.empty((n, h, w, ops.round_up(c, 8)), dtype=self.dtype, device=self.dev)
                    run_bwd(dy)
                    x.grads.append((dy, False))
                if affine:
                    self.add_param_grad(norm.weight, bstats[0, :, 1].contiguous())
                    self.add_param_grad(norm.bias, bstats[0, :, 0].contiguous())
            self.back.append(backward)
        return outv

    # ---------------------------------------------------------------- pooling / resampling / gates
    def avgpool2(self, x, out=None, halo=0, halo_kind=None):
        n, h, w, _ = x.t.shape
        outv = out if out is not None else self.new_val(n, h // 2, w // 2, x.c, halo, halo_kind)
        if out is None and halo and halo_kind != 'zero':
            raise NotImplementedError("avgpool2 writes the interior only (zero halo)")
        ops.avgpool2_fwd(x.t, outv.t)
        self._fill_stats(outv)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                dx = torch.empty(tuple(x.t.shape), dtype=self.dtype, device=self.dev)
                ops.avgpool2_bwd(g, dx)
                self.add_grad(x, dx)
            self.back.append(backward)
        return outv

    def bilinear2x(self, x, halo=0, halo_kind=None):
        n, h, w, _ = x.t.shape
        outv = self.new_val(n, 2 * h, 2 * w, x.c, halo, halo_kind)
        ops.bilinear2x_fwd(x.t, outv.t)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                dx = torch.empty(tuple(x.t.shape), dtype=self.dtype, device=self.dev)
                ops.bilinear2x_bwd(g, dx)
                self.add_grad(x, dx)
            self.back.append(backward)
        return outv

    def gate(self, base, s, att, halo=0, halo_kind=None, out=None):
        """out = [base +] sigmoid(mean_hw(att)) * s  (nn.AdaptiveAvgPool2d(1) + nn.Sigmoid + torch.mul [+ add],
        networks5_ds.py:641-649 and :696-700)."""
        n, h, w, _ = s.t.shape
        c = s.c
        na, ha, wa, _ = att.t.shape
        assert att.c == c and na == n
        sums = self.arena.take((n, c, 2))
        ops.channel_stats(att.t, c, True, sums)
        inv = 1.0 / float(ha * wa)
        outv = out if out is not None else self.new_val(n, h, w, c, halo, halo_kind)
        ops.gate_fwd(base.t if base is not None else None, s.t, sums, c, inv, outv.t)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                ds = torch.empty(tuple(s.t.shape), dtype=self.dtype, device=self.dev)
                dsum = self.arena.take((n, c))
                ops.gate_bwd(g, s.t, sums, c, inv, ds, dsum)
                self.add_grad(s, ds)
                if base is not None:
                    self.add_grad(base, g)
                dt = torch.empty(tuple(att.t.shape), dtype=self.dtype, device=self.dev)
                ops.gate_bcast(dsum, sums, c, inv, dt)
                self.add_grad(att, dt)
            self.back.append(backward)
        return outv

    def prelu(self, x, prelu_module, halo=0, halo_kind=None, out=None):
        if prelu_module.weight.numel() != 1:
            raise NotImplementedError("per-channel PReLU")
        if out is None and halo and halo_kind != 'zero':
            raise NotImplementedError("prelu writes the interior only (zero halo)")
        n, h, w, _ = x.t.shape
        outv = out if out is not None else self.new_val(n, h, w, x.c, halo, halo_kind)
        ops.prelu_fwd(x.t, prelu_module.weight, outv.t)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                dx = torch.empty(tuple(x.t.shape), dtype=self.dtype, device=self.dev)
                want = self.wants(prelu_module.weight)
                ds = torch.zeros((1,), dtype=torch.float32, device=self.dev) if want else None
                ops.prelu_bwd(x.t, g, prelu_module.weight, dx, ds)
                self.add_grad(x, dx)
                if want:
                    self.add_param_grad(prelu_module.weight, ds)
            self.back.append(backward)
        return outv

    def scale(self, x, alpha, out=None):
        """out = alpha * x (the down-weighted skip connections of models/encoder_decoder.py:196-205)."""
        n, h, w, _ = x.t.shape
        outv = out if out is not None else self.new_val(n, h, w, x.c)
        ops.scale(x.t, float(alpha), outv.t)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                dx = torch.empty(tuple(x.t.shape), dtype=self.dtype, device=self.dev)
                ops.scale(g, float(alpha), dx)
                self.add_grad(x, dx)
            self.back.append(backward)
        return outv

    def nearest2x(self, x, out=None):
        """nn.Upsample(scale_factor=2, mode='nearest') (models/encoder_decoder.py:193)."""
        n, h, w, _ = x.t.shape
        outv = out if out is not None else self.new_val(n, 2 * h, 2 * w, x.c)
        ops.nearest2x_fwd(x.t, outv.t)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                dx = torch.empty(tuple(x.t.shape), dtype=self.dtype, device=self.dev)
                ops.nearest2x_bwd(g, dx)
                self.add_grad(x, dx)
            self.back.append(backward)
        return outv

    def tanh(self, x):
        """nn.Tanh() on a value that stays inside the network (the intermediate output blocks of
        models/encoder_decoder.py:103-117 feed the next decoder level)."""
        n, h, w, _ = x.t.shape
        outv = self.new_val(n, h, w, x.c)
        ops.tanh_fwd(x.t, outv.t)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                dx = torch.empty(tuple(x.t.shape), dtype=self.dtype, device=self.dev)
                ops.tanh_bwd(outv.t, g, dx)
                self.add_grad(x, dx)
            self.back.append(backward)
        return outv

    def dropout(self, x, p):
        """In-place dropout (training mode only); the mask is regenerated from the seed in the backward."""
        if not self.training or p <= 0.0:
            return x
        # the seed lives in device memory: a captured training step draws a new mask at every replay
        seed = ops.dropout_seed(self.dev)
        ops.dropout_dev(x.t, x.t, seed, p)
        if self.record:
            def backward():
                g = self.total_grad(x)
                if g is None:
                    return
                dx = torch.empty(tuple(x.t.shape), dtype=self.dtype, device=self.dev)
                ops.dropout_dev(g, dx, seed, p)
                x.grads[:] = [(dx, False)]
                x.folded_parent = True
            self.back.append(backward)
        return x

    def add(self, a, b, halo=0, halo_kind=None):
        n, h, w, _ = a.t.shape
        outv = self.new_val(n, h, w, a.c, halo, halo_kind)
        ops.add(a.t, b.t, outv.t)
        if self.record:
            def backward():
                g = self.total_grad(outv)
                if g is None:
                    return
                self.add_grad(a, g)
                self.add_grad(b, g)
            self.back.append(backward)
        return outv

    # ---------------------------------------------------------------- replay
    def run_backward(self):
        for fn in reversed(self.back):
            fn()
        self.back = []


class GraphFunction(torch.autograd.Function):
    """One autograd node per network call: forward runs ``body(tape, *inputs)`` which returns
    (list of fp32 output tensors, list of output slots, list of input Vals)."""

    @staticmethod
    def forward(ctx, body, module, n_inputs, *tensors):
        inputs = tensors[:n_inputs]
        params = tensors[n_inputs:]
        dev = inputs[0].device
        record = any(ctx.needs_input_grad[3:])
        tape = Tape(module.training, dev, record)
        tape.input_wants = [bool(f) for f in ctx.needs_input_grad[3:3 + n_inputs]]
        outs, slots, in_vals = body(tape, *[t.detach() for t in inputs])
        ctx.tape, ctx.slots, ctx.in_vals = tape, slots, in_vals
        ctx.in_shapes = [tuple(t.shape) for t in inputs]
        ctx.params = params
        ctx.n_inputs = n_inputs
        nd = [o for o, s in zip(outs, slots) if not s['differentiable']]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        tape = ctx.tape
        if tape is None:
            raise RuntimeError("cdb200: backward called twice on the same network call")
        n_in = ctx.n_inputs
        tape.needs = {p: bool(ctx.needs_input_grad[3 + n_in + i]) for i, p in enumerate(ctx.params)}
        for slot, g in zip(ctx.slots, gouts):
            tape.seed_output_grad(slot, g)
        tape.run_backward()
        gin = []
        for i, v in enumerate(ctx.in_vals):
            if v is not None and ctx.needs_input_grad[3 + i]:
                gin.append(tape.input_grad(v, ctx.in_shapes[i]))
            else:
                gin.append(None)
        gp = [tape.param_grads.get(p) if tape.needs[p] else None for p in ctx.params]
        ctx.tape = None
        return (None, None, None) + tuple(gin) + tuple(gp)


def run(module, body, inputs):
    """Runs `body` on the tape with autograd support. Returns the tuple of output tensors."""
    params = [p for p in module.parameters()]
    seen, uniq = set(), []
    for p in params:
        if id(p) not in seen:
            seen.add(id(p))
            uniq.append(p)
    if torch.is_grad_enabled() and (any(t.requires_grad for t in inputs) or any(p.requires_grad for p in uniq)):
        return GraphFunction.apply(body, module, len(inputs), *inputs, *uniq)
    tape = Tape(module.training, inputs[0].device, False)
    tape.input_wants = [False] * len(inputs)
    outs, _, _ = body(tape, *[t.detach() for t in inputs])
    return tuple(outs)
