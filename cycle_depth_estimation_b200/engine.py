"""Fused execution of the reference's convolutional networks on the B200 kernels.

The drop-in modules of ``networks.py`` keep the reference's module tree (and therefore its
state_dict), but their ``forward`` does not run the tree: it is compiled once into a list of fused
*stages*  [pad] -> conv -> [norm] -> [activation] -> [+ residual]  and executed by one
``torch.autograd.Function`` per network call.  Between stages activations live as NHWC bf16 buffers
whose reflect halo is written by the producing kernel, so padding, normalisation, activation and the
residual add never exist as separate passes.  The backward pass is hand-written: halo fold + activation
mask + norm backward (one fused kernel pair), tcgen05 wgrad and tcgen05 dgrad per stage.

Reference structure being executed: models/networks.py:145-191 (ResnetGenerator), :195-236
(ResnetBlock), :320-364 (NLayerDiscriminator), :367-389 (PixelDiscriminator).
"""
import os as _os

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, NORM_BATCH, NORM_INSTANCE, NORM_NONE

BF16 = torch.bfloat16


class Stage:
    """One fused conv stage. Values are numbered; value 0 is the network input."""

    def __init__(self):
        self.conv = None          # nn.Conv2d / nn.ConvTranspose2d (parameter holder)
        self.transposed = False
        self.reflect = 0          # ReflectionPad2d in front of the conv
        self.norm = None          # nn.InstanceNorm2d / nn.BatchNorm2d / None
        self.act = ACT_NONE
        self.slope = 0.0
        self.src = 0              # input value id
        self.dst = 0              # output value id
        self.res = None           # value id added after norm (ResnetBlock)

    def __repr__(self):
        return "Stage(%s k%d s%d reflect%d norm=%s act=%d src=%d dst=%d res=%s)" % (
            "convT" if self.transposed else "conv", self.conv.kernel_size[0], self.conv.stride[0], self.reflect,
            type(self.norm).__name__ if self.norm is not None else None, self.act, self.src, self.dst, self.res)


def _act_of(m):
    if isinstance(m, nn.ReLU):
        return ACT_RELU, 0.0
    if isinstance(m, nn.LeakyReLU):
        return ACT_LEAKY, float(m.negative_slope)
    if isinstance(m, nn.Tanh):
        return ACT_TANH, 0.0
    if isinstance(m, nn.Sigmoid):
        return ACT_SIGMOID, 0.0
    return None


def compile_chain(modules, stages=None, cur=0, next_id=None):
    """Compiles a flat list of modules (nn.Sequential children; residual blocks expose
    ``conv_block``) into stages. Returns (stages, id of the final value)."""
    stages = [] if stages is None else stages
    next_id = [1] if next_id is None else next_id
    pending_reflect = 0
    i = 0
    mods = list(modules)
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.ReflectionPad2d):
            p = m.padding
            if not (p[0] == p[1] == p[2] == p[3]):
                raise NotImplementedError("asymmetric ReflectionPad2d")
            pending_reflect = int(p[0])
            i += 1
            continue
        if hasattr(m, "conv_block") and isinstance(m.conv_block, nn.Sequential):
            # residual block: out = x + conv_block(x)
            block_in = cur
            _, cur = compile_chain(list(m.conv_block.children()), stages, cur, next_id)
            last = stages[-1]
            if last.act != ACT_NONE:
                raise NotImplementedError("residual add after an activation")
            last.res = block_in
            i += 1
            continue
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            st = Stage()
            st.conv = m
            st.transposed = isinstance(m, nn.ConvTranspose2d)
            st.reflect = pending_reflect
            pending_reflect = 0
            st.src = cur
            st.dst = next_id[0]
            next_id[0] += 1
            cur = st.dst
            i += 1
            if i < len(mods) and isinstance(mods[i], (nn.InstanceNorm2d, nn.BatchNorm2d)):
                st.norm = mods[i]
                i += 1
            if i < len(mods) and _act_of(mods[i]) is not None:
                st.act, st.slope = _act_of(mods[i])
                i += 1
            if i < len(mods) and isinstance(mods[i], nn.Dropout):
                raise NotImplementedError("Dropout inside a fused chain")
            _check_conv(st)
            stages.append(st)
            continue
        raise NotImplementedError("cdb200 engine: unsupported module %s" % type(m).__name__)
    if pending_reflect:
        raise NotImplementedError("trailing ReflectionPad2d")
    return stages, cur


def _check_conv(st):
    c = st.conv
    if c.groups != 1:
        raise NotImplementedError("grouped convolution")
    if c.kernel_size[0] != c.kernel_size[1] or c.stride[0] != c.stride[1] or c.padding[0] != c.padding[1] \
            or c.dilation[0] != c.dilation[1]:
        raise NotImplementedError("non-square convolution geometry")
    if c.stride[0] not in (1, 2):
        raise NotImplementedError("stride %d" % c.stride[0])
    if getattr(c, "padding_mode", "zeros") != "zeros":
        raise NotImplementedError("padding_mode %s" % c.padding_mode)
    if st.reflect and (c.padding[0] != 0 or st.transposed):
        raise NotImplementedError("ReflectionPad2d followed by a padded / transposed convolution")
    if st.norm is not None and isinstance(st.norm, nn.InstanceNorm2d) and (st.norm.affine or st.norm.track_running_stats):
        raise NotImplementedError("InstanceNorm2d with affine / running stats")


# -------------------------------------------------------------------------------------------------
# packed weight cache (invalidated by the parameter's version counter, bumped by optimizer steps)
# -------------------------------------------------------------------------------------------------
_pack_epoch = [0]


def invalidate_packed_weights():
    """Drops every cached bf16 weight packing. The cache notices in-place updates that bump the tensor
    version counter (optimizers, load_state_dict, copy_) and FusedAdam steps; writes through ``.data``
    are invisible to it, so code that does that (init_weights does) calls this."""
    _pack_epoch[0] += 1


class _PackCache:
    """bf16 GEMM-operand copies of the fp32 filters, stored ON the parameter object (so they die with
    it) and validated against its storage pointer and version counters."""

    def get(self, weight, rows_are_dim0, rowpack, flipped=False):
        """flipped: pack W[..., R-1-r, S-1-s] (row-packed data gradients, where the kernel cannot reverse taps)."""
        store = weight.__dict__.setdefault('_cdb_packed', {})
        key = (rows_are_dim0, rowpack, flipped)
        ver = (weight.data_ptr(), weight._version, getattr(weight, '_cdb_version', 0), _pack_epoch[0])
        hit = store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        src = weight.detach().flip(2, 3) if flipped else weight.detach()
        packed = ops.pack_conv_weight(src.contiguous(), rows_are_dim0, rowpack,
                                      out=hit[1][0] if hit is not None else None)
        store[key] = (ver, packed)
        return packed


    def get_folded(self, weight, dgrad):
        """Few-output-channel layers: the S filter columns folded into the GEMM N dimension.
        forward: W2[s*O + o, i, r, 0] = W[o, i, r, s];  data gradient: W2[s*I + i, o, r, 0] = W[o, i, R-1-r, S-1-s]."""
        store = weight.__dict__.setdefault('_cdb_packed', {})
        key = ('fold', dgrad)
        ver = (weight.data_ptr(), weight._version, getattr(weight, '_cdb_version', 0), _pack_epoch[0])
        hit = store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        w = weight.detach()
        o, i, r, sdim = w.shape
        if dgrad:
            w2 = w.flip(2, 3).permute(3, 1, 0, 2).reshape(sdim * i, o, r, 1)
        else:
            w2 = w.permute(3, 0, 1, 2).reshape(sdim * o, i, r, 1)
        packed = ops.pack_conv_weight(w2.contiguous(), True, 0, out=hit[1][0] if hit is not None else None)
        store[key] = (ver, packed)
        return packed


    def get_toeplitz(self, weight, rows_are_dim0, flip):
        """Toeplitz operand of an image layer (ops.pack_toeplitz_weight): (packed, rows_pad)."""
        store = weight.__dict__.setdefault('_cdb_packed', {})
        key = ('tz', rows_are_dim0, flip)
        ver = (weight.data_ptr(), weight._version, getattr(weight, '_cdb_version', 0), _pack_epoch[0])
        hit = store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        packed = ops.pack_toeplitz_weight(weight.detach().contiguous(), rows_are_dim0, flip,
                                          out=hit[1][0] if hit is not None else None)
        store[key] = (ver, packed)
        return packed

    def get_tf32(self, weight, rows_are_dim0, cs, x3, flipped=False):
        """fp32 (TF32-rounded) GEMM operand of the fp32-storage network modes.  cs: stored channels of the
        activation the convolution contracts over.  x3 (error-compensated mode): the contraction dimension holds
        [w_hi | w_hi | w_lo], each block zero-padded to cs channels, matching the [x_hi | x_lo | x_hi] operand
        ops.split_tf32(x, 0) produces."""
        store = weight.__dict__.setdefault('_cdb_packed', {})
        key = ('tf32', rows_are_dim0, cs, x3, flipped)
        ver = (weight.data_ptr(), weight._version, getattr(weight, '_cdb_version', 0), _pack_epoch[0])
        hit = store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        w = weight.detach().flip(2, 3) if flipped else weight.detach()
        if x3:
            kd = 1 if rows_are_dim0 else 0
            hi = ops.round_tf32_(w.clone(memory_format=torch.contiguous_format))
            lo = w - hi                       # exact in fp32; rounded to TF32 by the packing kernel
            shape = list(w.shape)
            shape[kd] = 3 * cs
            w3 = torch.zeros(shape, dtype=torch.float32, device=w.device)
            k = w.shape[kd]
            for blk, part in enumerate((hi, hi, lo)):
                w3.narrow(kd, blk * cs, k).copy_(part)
            w = w3
        packed = ops.pack_conv_weight_tf32(w.contiguous(), rows_are_dim0)
        store[key] = (ver, packed)
        return packed

    def adam_targets(self, weight):
        """Packed bf16 layouts of `weight` that the fused Adam kernel refreshes itself: at most two plain
        (unflipped, unfolded) ones.  Returns [(packed tensor, rows_are_dim0, rowpack)]."""
        store = weight.__dict__.get('_cdb_packed')
        if not store or weight.dim() != 4 or not weight.is_contiguous():
            return []
        out = []
        for key, (_, packed) in store.items():
            if len(key) == 3 and isinstance(key[0], bool) and not key[2]:
                out.append((packed[0], key[0], key[1]))
                if len(out) == 2:
                    break
        return out

    def stamp(self, weight, packs):
        """Marks the layouts in `packs` (just rewritten by the Adam kernel) as current for the weight's new version."""
        if not packs:
            return
        store = weight.__dict__['_cdb_packed']
        ver = (weight.data_ptr(), weight._version, getattr(weight, '_cdb_version', 0), _pack_epoch[0])
        bufs = {id(b) for b, _, _ in packs}
        for key, (old_ver, packed) in list(store.items()):
            if id(packed[0]) in bufs:
                store[key] = (ver, packed)

    def refresh(self, params):
        """After an optimizer step: re-packs every cached plain layout of `params` with ONE multi-tensor launch
        and stamps the cache entries with the parameters' new versions (flipped / folded layouts of the two 7x7
        layers are left to the lazy path)."""
        items, stamps = [], []
        for w in params:
            store = w.__dict__.get('_cdb_packed')
            if not store or w.dim() != 4:
                continue
            ver = (w.data_ptr(), w._version, getattr(w, '_cdb_version', 0), _pack_epoch[0])
            for key, (old_ver, packed) in store.items():
                if len(key) == 3 and isinstance(key[0], bool) and not key[2] and old_ver != ver and w.is_contiguous():
                    items.append((w.detach(), key[0], key[1], packed[0]))
                    stamps.append((store, key, ver, packed))
        ops.pack_conv_weights_multi(items)
        for store, key, ver, packed in stamps:
            store[key] = (ver, packed)


_pack_cache = _PackCache()


def _foldable(k, channels, conv):
    """S*channels output columns fit one narrow MMA tile: worth folding (7x7 with 3 channels -> N = 21)."""
    return k > 1 and k * channels <= 32 and conv.stride[0] == 1 and conv.dilation[0] == 1


def _out_hw(st, h, w):
    c = st.conv
    k, s, p, d = c.kernel_size[0], c.stride[0], c.padding[0], c.dilation[0]
    if st.transposed:
        op = c.output_padding[0]
        return (h - 1) * s - 2 * p + d * (k - 1) + op + 1, (w - 1) * s - 2 * p + d * (k - 1) + op + 1
    hp, wp = h + 2 * st.reflect, w + 2 * st.reflect
    return (hp + 2 * p - d * (k - 1) - 1) // s + 1, (wp + 2 * p - d * (k - 1) - 1) // s + 1


class Plan:
    """Compiled network: stages + per-value metadata derived from the consumers."""

    def __init__(self, stages, final):
        self.stages = stages
        self.final = final
        self.n_values = final + 1
        # halo each value must carry = reflect pad of the conv that reads it
        self.halo = [0] * self.n_values
        self.readers = [[] for _ in range(self.n_values)]
        for idx, st in enumerate(stages):
            self.readers[st.src].append(idx)
            if st.reflect:
                self.halo[st.src] = max(self.halo[st.src], st.reflect)
        # zero halo materialised for stride-1 zero-padded direct convolutions (lets them take the flat kernel)
        self.zero_halo = [0] * self.n_values
        for idx, st in enumerate(stages):
            c = st.conv
            if (st.src != 0 and not st.transposed and not st.reflect and c.stride[0] == 1 and c.padding[0] > 0):
                self.zero_halo[st.src] = c.padding[0]
        for v in range(1, self.n_values):
            rs = self.readers[v]
            if len(rs) > 1:
                raise NotImplementedError("value read by several convolutions")
        first = stages[0]
        if first.src != 0 or first.transposed:
            raise NotImplementedError("network must start with a convolution of its input")
        last = stages[-1]
        if last.dst != final or last.norm is not None or last.res is not None:
            raise NotImplementedError("network must end with a plain convolution (+ activation)")
        self.params = []
        for st in stages:
            self.params.append(st.conv.weight)
            if st.conv.bias is not None:
                self.params.append(st.conv.bias)
            if st.norm is not None and getattr(st.norm, "affine", False):
                self.params.append(st.norm.weight)
                self.params.append(st.norm.bias)


def _rowpack_for(st, cin):
    k = st.conv.kernel_size[0]
    if st.transposed or st.conv.dilation[0] != 1:
        return 0
    if cin <= 8 and k <= 8:
        return 8
    if cin <= 16 and k <= 4:
        return 16
    return 0


def _toeplitz_first(st, cin):
    """The first stage is an image layer the Toeplitz kernels take (csrc/conv_toeplitz.cu): stride 1, no dilation,
    <= 8 input channels, filter <= 8 x 8, <= 128 output channels — c7s1-64 of the generators."""
    c = st.conv
    return (not st.transposed and c.stride[0] == 1 and c.dilation[0] == 1 and cin <= 8 and c.kernel_size[0] <= 8
            and c.out_channels <= 128 and not _os.environ.get("CDB_NO_TOEPLITZ"))


def _toeplitz_last(conv, transposed, cs):
    """The data / weight gradient of a stride-1 layer with <= 8 output channels (c7s1-3): its 'image' is dy."""
    return (not transposed and cs == 8 and conv.dilation[0] == 1 and conv.kernel_size[0] <= 8 and conv.in_channels <= 128
            and not _os.environ.get("CDB_NO_TOEPLITZ"))


def _norm_kind(st):
    if st.norm is None:
        return NORM_NONE
    return NORM_INSTANCE if isinstance(st.norm, nn.InstanceNorm2d) else NORM_BATCH


def _flat_stage(plan, st, is_first, rowpack):
    """Stride-1 direct convolution whose input buffer carries its padding -> flat kernel."""
    if st.transposed or st.conv.stride[0] != 1 or (is_first and rowpack):
        return False
    if st.reflect:
        return True
    return st.conv.padding[0] == 0 or (st.src != 0 and plan.zero_halo[st.src] == st.conv.padding[0])


class _Run:
    """State of one forward call kept for its backward."""
    pass


def _geom_fwd(st, materialised_pad, rowpack=0):
    c = st.conv
    pad = 0 if materialised_pad else c.padding[0]
    return ops.geom(c.kernel_size[0], c.kernel_size[1], c.stride[0], pad, pad, c.dilation[0], st.transposed, rowpack)


def _geom_dgrad(st, materialised_pad):
    c = st.conv
    pad = 0 if materialised_pad else c.padding[0]
    return ops.geom(c.kernel_size[0], c.kernel_size[1], c.stride[0], pad, pad, c.dilation[0], not st.transposed, 0)


def forward(plan, x, training, need_input_grad, needs_param_grad):
    """Runs the plan. x: fp32 NCHW CUDA tensor. Returns (output fp32 NCHW, run state)."""
    if not x.is_cuda:
        raise RuntimeError("cdb200 networks run on CUDA tensors only (no CPU path)")
    if x.dtype != torch.float32:
        raise TypeError("fp32 NCHW input expected")
    dev = x.device
    n, cin, h, w = x.shape
    run = _Run()
    run.x_shape = tuple(x.shape)
    run.vals = [None] * plan.n_values    # full buffers (with halo)
    run.inner = [None] * plan.n_values   # interior views
    run.dims = [None] * plan.n_values    # (h, w, c)
    run.y = [None] * len(plan.stages)
    run.stats = [None] * len(plan.stages)
    run.training = training
    run.out = None
    arena = ops.ZeroArena(dev)

    # ---- value 0: the image, converted to NHWC bf16 with the first conv's padding materialised
    st0 = plan.stages[0]
    rp = _rowpack_for(st0, cin)
    run.rowpack = rp
    run.toeplitz0 = bool(rp) and rp == 8 and _toeplitz_first(st0, cin)
    conv0 = st0.conv
    pad0 = st0.reflect if st0.reflect else conv0.padding[0]
    if rp:
        ho, wo = _out_hw(st0, h, w)
        span = 64 // rp
        wneed = max(w + 2 * pad0, conv0.stride[0] * (wo - 1) + span)
        buf = torch.zeros((n, h + 2 * pad0, wneed, rp), dtype=BF16, device=dev)
        inner = buf[:, pad0:pad0 + h, pad0:pad0 + w, :]
        ops.nchw_to_nhwc(x, inner, pad=st0.reflect)
        run.vals[0] = buf
        run.inner[0] = inner
        run.in_pad = pad0
    else:
        cs = ops.round_up(cin, 8)
        pr = st0.reflect
        buf = torch.empty((n, h + 2 * pr, w + 2 * pr, cs), dtype=BF16, device=dev)
        inner = buf[:, pr:pr + h, pr:pr + w, :]
        ops.nchw_to_nhwc(x, inner, pad=pr)
        run.vals[0] = buf
        run.inner[0] = inner
        run.in_pad = pr
    run.dims[0] = (h, w, cin)

    for idx, st in enumerate(plan.stages):
        conv = st.conv
        hi, wi, ci = run.dims[st.src]
        ho, wo = _out_hw(st, hi, wi)
        co = conv.out_channels
        cs = ops.round_up(co, 8)
        is_first = idx == 0
        is_last = idx == len(plan.stages) - 1
        rowpack = rp if is_first else 0
        zero_mat = st.src != 0 and plan.zero_halo[st.src] > 0
        materialised = bool(st.reflect) or bool(is_first and rp) or zero_mat
        flat = _flat_stage(plan, st, is_first, rowpack)
        g = _geom_fwd(st, materialised, rowpack)
        wp, rows_pad, kpad = _pack_cache.get(conv.weight, not st.transposed, rowpack)
        xin = run.vals[st.src] if materialised else run.inner[st.src]
        nk = _norm_kind(st)
        if is_last:
            out = torch.empty((n, co, ho, wo), dtype=torch.float32, device=dev)
            k = conv.kernel_size[0]
            if flat and xin.is_contiguous() and _foldable(k, co, conv):
                # c7s1-3: R x 1 convolution with S*Cout folded output channels, then the shifted row sum
                w2, rows2, kpad2 = _pack_cache.get_folded(conv.weight, False)
                t = torch.empty((n, ho, xin.shape[2], ops.round_up(k * co, 8)), dtype=torch.float32, device=dev)
                ops.conv2d_fwd(ops.geom(k, 1), xin, w2, rows2, kpad2, ops.out_view_nhwc(t, k * co))
                ops.shift_add_nchw(t, k, co, conv.bias, st.act, st.slope, out)
            else:
                ops.conv2d_fwd(g, xin, wp, rows_pad, kpad, ops.out_view_nchw(out), conv.bias, st.act, st.slope)
            run.out = out
            run.dims[st.dst] = (ho, wo, co)
            continue
        halo = plan.halo[st.dst]
        zh = plan.zero_halo[st.dst]
        if zh:
            if halo:
                raise NotImplementedError("value with both a reflect and a zero halo")
            dbuf = ops.empty_zero_halo(n, ho, wo, cs, zh, 0, dev)
            dinner = dbuf[:, zh:zh + ho, zh:zh + wo, :]
        else:
            dbuf = torch.empty((n, ho + 2 * halo, wo + 2 * halo, cs), dtype=BF16, device=dev)
            dinner = dbuf[:, halo:halo + ho, halo:halo + wo, :]
        run.vals[st.dst] = dbuf
        run.inner[st.dst] = dinner
        run.dims[st.dst] = (ho, wo, co)
        tz = is_first and run.toeplitz0
        if tz:
            twp, trows = _pack_cache.get_toeplitz(conv.weight, True, False)
        if nk == NORM_NONE:
            if halo:
                raise NotImplementedError("reflect padding after a stage without normalisation")
            if st.res is not None:
                raise NotImplementedError("residual add without normalisation")
            if tz:
                ops.conv2d_toeplitz_fwd(xin, twp, trows, conv.kernel_size[0], conv.kernel_size[1],
                                        ops.out_view_nhwc(dinner, co), conv.bias, st.act, st.slope)
            else:
                ops.conv2d_fwd(g, xin, wp, rows_pad, kpad, ops.out_view_nhwc(dinner, co), conv.bias, st.act, st.slope)
            continue
        # conv -> raw y (+ fused per-channel sums) -> norm/act/residual/halo
        if flat or tz:
            y = ops.alloc_flat_output(n, ho, wo, xin.shape[2], cs, dev)  # pitched: TMA-store epilogue
        else:
            y = torch.empty((n, ho, wo, cs), dtype=BF16, device=dev)
        use_running = nk == NORM_BATCH and not training and st.norm.track_running_stats
        stats = None
        # A bias in front of a normalisation cancels exactly (InstanceNorm affine=False /
        # BatchNorm in training mode); it is skipped and its gradient is zero (SURVEY B-4).
        bias = conv.bias if use_running else None
        if not use_running:
            groups = n if nk == NORM_INSTANCE else 1
            stats = arena.take((groups, co, 2))
        if tz:
            ops.conv2d_toeplitz_fwd(xin, twp, trows, conv.kernel_size[0], conv.kernel_size[1], ops.out_view_nhwc(y, co),
                                    bias, ACT_NONE, 0.0, stats if nk == NORM_INSTANCE else None)
            if nk != NORM_INSTANCE and not use_running:
                ops.channel_stats(y, co, False, stats)
        elif nk == NORM_INSTANCE:
            ops.conv2d_fwd(g, xin, wp, rows_pad, kpad, ops.out_view_nhwc(y, co), bias, ACT_NONE, 0.0, stats)
        else:
            ops.conv2d_fwd(g, xin, wp, rows_pad, kpad, ops.out_view_nhwc(y, co), bias, ACT_NONE, 0.0, None)
            if not use_running:
                ops.channel_stats(y, co, False, stats)
        affine = getattr(st.norm, "affine", False)
        bn_world = ops.bn_world() if (nk == NORM_BATCH and not use_running) else 1
        if bn_world > 1:
            ops.bn_all_reduce(stats)      # batch statistics over the shards of all ranks
        desc = ops.norm_desc(nk, st.act, st.slope, float(st.norm.eps), co, halo, stats,
                             st.norm.weight if affine else None, st.norm.bias if affine else None,
                             st.norm.running_mean if nk == NORM_BATCH else None,
                             st.norm.running_var if nk == NORM_BATCH else None,
                             use_running=use_running,
                             update_running=(nk == NORM_BATCH and training and st.norm.track_running_stats),
                             momentum=float(st.norm.momentum) if getattr(st.norm, "momentum", None) is not None else 0.1,
                             conv_bias=conv.bias if bias is None else None, count_scale=bn_world)
        if nk == NORM_BATCH and training and st.norm.track_running_stats and st.norm.num_batches_tracked is not None:
            st.norm.num_batches_tracked += 1
        res = run.inner[st.res] if st.res is not None else None
        ops.norm_act_fwd(desc, y, dinner, res)
        run.y[idx] = y
        run.stats[idx] = stats
    return run.out, run


DEBUG_RECORD = None  # set to a dict by tools/debug_bwd.py to capture per-stage gradients


class _SideStream:
    """Runs the weight gradients of a network call's backward on a second CUDA stream.  ``fork`` makes the side
    stream wait for everything the current stream has issued so far (the gradient dy of the stage) and keeps the
    tensors the side-stream kernels read alive until ``join`` — the caching allocator may otherwise hand their memory
    to a later allocation of the main stream while the side stream is still reading it.  Works inside a CUDA-graph
    capture (the event dependencies become graph edges: the weight gradients are parallel branches of the graph)."""
    _streams = {}

    def __init__(self, device, enabled):
        self.enabled = enabled
        self.keep = []
        self.used = False
        if enabled:
            # one companion stream per (device, calling stream): two network calls running on two streams (the two
            # discriminators of a CycleGAN update) keep independent weight-gradient branches
            key = (device.index if device.index is not None else torch.cuda.current_device(),
                   torch.cuda.current_stream().cuda_stream)
            st = _SideStream._streams.get(key)
            if st is None:
                st = torch.cuda.Stream(device=device)
                _SideStream._streams[key] = st
            self.stream = st
            self.ctx = None

    def fork(self, *tensors):
        if not self.enabled:
            return
        self.keep.extend(t for t in tensors if t is not None)
        self.stream.wait_stream(torch.cuda.current_stream())
        self.used = True

    def __enter__(self):
        if self.enabled:
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.enabled:
            self.ctx.__exit__(*exc)
            self.ctx = None
        return False

    def join(self):
        if self.enabled and self.used:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep = []


_WGRAD_SIDE_STREAM = [not _os.environ.get("CDB_NO_WGRAD_STREAM")]


def backward(plan, run, gout, need_input_grad, needs_param_grad):
    """Hand-written backward. Returns (grad_input or None, {param: grad})."""
    dev = gout.device
    side = _SideStream(dev, _WGRAD_SIDE_STREAM[0])
    n = run.x_shape[0]
    grads = {}
    stages = plan.stages
    # gradient contributions per value: padded conv-dgrad result and skip (residual) gradient
    arena = ops.ZeroArena(dev)
    dpad = [None] * plan.n_values   # full buffer incl. halo (gradient w.r.t. the padded value)
    dskip = [None] * plan.n_values
    gx = None
    gout = gout.contiguous()
    for idx in range(len(stages) - 1, -1, -1):
        st = stages[idx]
        conv = st.conv
        is_first = idx == 0
        is_last = idx == len(stages) - 1
        ho, wo, co = run.dims[st.dst]
        hi, wi, ci = run.dims[st.src]
        cs = ops.round_up(co, 8)
        nk = _norm_kind(st)
        rowpack = run.rowpack if is_first else 0
        zero_mat = st.src != 0 and plan.zero_halo[st.src] > 0
        materialised = bool(st.reflect) or bool(is_first and rowpack) or zero_mat
        flat = _flat_stage(plan, st, is_first, rowpack)
        want_w = needs_param_grad.get(conv.weight, False)
        want_b = conv.bias is not None and needs_param_grad.get(conv.bias, False)
        affine = st.norm is not None and getattr(st.norm, "affine", False)
        want_dx = (not is_first) or need_input_grad
        # ---- dY: gradient w.r.t. the raw convolution output. For flat stages it lives inside a zero halo
        # of (k-1)*dil pixels, so that the data gradient is again a flat (flipped) convolution.
        flat_dgrad = flat and want_dx and not is_first
        fold_dgrad = (is_first and want_dx and bool(st.reflect) and not st.transposed
                      and _foldable(conv.kernel_size[0], run.x_shape[1], conv))
        if flat_dgrad or fold_dgrad:
            hz = (conv.kernel_size[0] - 1) * conv.dilation[0]
            slack = 64 // cs if cs <= 16 else 0   # room for the row-packed view used by the few-channel wgrad
            dyp = ops.empty_zero_halo(n, ho, wo, cs, hz, slack, dev)
            dy = dyp[:, hz:hz + ho, hz:hz + wo, :]
        else:
            dy = torch.empty((n, ho, wo, cs), dtype=BF16, device=dev)
        if is_last:
            ops.nchw_to_nhwc(gout, dy, pad=0, act_out=run.out if st.act != ACT_NONE else None, act=st.act,
                             slope=st.slope)
            if want_b:
                db = torch.empty((co,), dtype=torch.float32, device=dev)
                ops.bias_grad_nchw(gout, run.out if st.act != ACT_NONE else None, st.act, st.slope, db)
                grads[conv.bias] = db
        else:
            halo = plan.halo[st.dst]
            off = halo + plan.zero_halo[st.dst]
            dout_full = dpad[st.dst]
            dout_inner = dout_full[:, off:off + ho, off:off + wo, :] if dout_full is not None else None
            use_running = nk == NORM_BATCH and not run.training and st.norm.track_running_stats
            groups = n if nk == NORM_INSTANCE else 1
            need_b = (nk != NORM_NONE and not use_running) or (nk == NORM_NONE and want_b) or affine
            bstats = arena.take((groups, co, 2)) if need_b else None
            bn_world = ops.bn_world() if (nk == NORM_BATCH and not use_running) else 1

            def desc_of(extra_flags):
                return ops.norm_desc(nk, st.act, st.slope, float(st.norm.eps) if st.norm is not None else 0.0, co, halo,
                                     run.stats[idx], st.norm.weight if affine else None,
                                     st.norm.bias if affine else None,
                                     st.norm.running_mean if nk == NORM_BATCH else None,
                                     st.norm.running_var if nk == NORM_BATCH else None, use_running=use_running,
                                     flags=extra_flags, count_scale=bn_world)
            yv = run.y[idx] if nk != NORM_NONE else run.inner[st.dst]
            gsum = None
            if st.res is not None:
                gsum = torch.empty((n, ho, wo, cs), dtype=BF16, device=dev)
            if bn_world > 1:
                ops.norm_act_bwd_synced(desc_of, yv, dy, dout_inner, dskip[st.dst], bstats, gsum, bn_world)
            else:
                ops.norm_act_bwd(desc_of(0), yv, dy, dout_inner, dskip[st.dst], bstats, gsum)
            if st.res is not None:
                if dskip[st.res] is not None:
                    raise NotImplementedError("value with two skip gradients")
                dskip[st.res] = gsum
            if nk == NORM_NONE and want_b:
                grads[conv.bias] = bstats[0, :, 0].contiguous()
            elif want_b:
                if use_running:
                    bs = arena.take((1, co, 2))
                    ops.channel_stats(dy, co, False, bs)
                    grads[conv.bias] = bs[0, :, 0].contiguous()
                else:
                    # exactly zero (a bias in front of a statistics-normalised layer): a slice of the zero arena, no
                    # fill kernel of its own (autograd takes the fresh view over as .grad without a copy)
                    grads[conv.bias] = arena.take(tuple(conv.bias.shape))
            if affine:
                if needs_param_grad.get(st.norm.weight, False):
                    grads[st.norm.weight] = bstats[0, :, 1].contiguous()
                if needs_param_grad.get(st.norm.bias, False):
                    grads[st.norm.bias] = bstats[0, :, 0].contiguous()
            dpad[st.dst] = None
            dskip[st.dst] = None
        if DEBUG_RECORD is not None:
            DEBUG_RECORD[('dy', idx)] = dy.clone()
        # ---- wgrad: on the side stream, concurrently with the data-gradient / norm-backward chain of the earlier
        # layers (it only has to be finished when the network call's backward returns)
        xin = run.vals[st.src] if materialised else run.inner[st.src]
        if want_w:
            side.fork(dy, dyp if (flat_dgrad or fold_dgrad) else None)
            with side:
                k = conv.kernel_size[0]
                if flat_dgrad and _toeplitz_last(conv, st.transposed, cs) and xin.is_contiguous():
                    # few output channels (the 7x7 c7s1-3 layer): the padded 64-channel input is iterated, the zero-haloed
                    # 3-channel dy is the shifted Toeplitz operand (taps reversed)
                    dw = torch.empty_like(conv.weight, memory_format=torch.contiguous_format)
                    ops.conv2d_toeplitz_wgrad(xin, dyp, k, k, dw, False, flip=True)
                elif is_first and run.toeplitz0:
                    dw = torch.empty_like(conv.weight, memory_format=torch.contiguous_format)
                    ops.conv2d_toeplitz_wgrad(dy, xin, k, k, dw, True)
                elif flat_dgrad and cs <= 16 and k * cs <= 64 and conv.dilation[0] == 1:
                    # few output channels (the 7x7 c7s1-3 layer): correlate the padded input with the zero-haloed
                    # dy, whose 8 channels x 8 pixels form one K block per filter row; the result comes out as
                    # [cin, cout, R-1-r, S-1-s]
                    tmp = torch.empty((ci, co, k, k), dtype=torch.float32, device=dev)
                    ops.conv2d_wgrad(ops.geom(k, k, 1, 0, 0, 1, True, cs), xin, dyp, tmp, False)
                    dw = tmp.flip(2, 3).permute(1, 0, 2, 3).contiguous()
                else:
                    dw = torch.empty_like(conv.weight, memory_format=torch.contiguous_format)
                    ops.conv2d_wgrad(_geom_fwd(st, materialised, rowpack), xin, dy, dw, False)
                grads[conv.weight] = dw
        # ---- dgrad
        if want_dx:
            gd = _geom_dgrad(st, materialised and not (is_first and rowpack and not st.reflect))
            wd, rows_pad, kpad = _pack_cache.get(conv.weight, st.transposed, 0)
            if is_first:
                cin = run.x_shape[1]
                if fold_dgrad:
                    p, k = st.reflect, conv.kernel_size[0]
                    w2, rows2, kpad2 = _pack_cache.get_folded(conv.weight, True)
                    tmp = torch.empty((n, cin, hi + 2 * p, wi + 2 * p), dtype=torch.float32, device=dev)
                    t = torch.empty((n, hi + 2 * p, dyp.shape[2], ops.round_up(k * cin, 8)), dtype=torch.float32,
                                    device=dev)
                    ops.conv2d_fwd(ops.geom(k, 1), dyp, w2, rows2, kpad2, ops.out_view_nhwc(t, k * cin))
                    ops.shift_add_nchw(t, k, cin, None, ACT_NONE, 0.0, tmp)
                    gx = torch.empty((n, cin, hi, wi), dtype=torch.float32, device=dev)
                    ops.reflect_fold_nchw(tmp, gx, p)
                elif st.reflect:
                    p = st.reflect
                    tmp = torch.empty((n, cin, hi + 2 * p, wi + 2 * p), dtype=torch.float32, device=dev)
                    ops.conv2d_fwd(gd, dy, wd, rows_pad, kpad, ops.out_view_nchw(tmp))
                    gx = torch.empty((n, cin, hi, wi), dtype=torch.float32, device=dev)
                    ops.reflect_fold_nchw(tmp, gx, p)
                else:
                    gx = torch.empty((n, cin, hi, wi), dtype=torch.float32, device=dev)
                    ops.conv2d_fwd(gd, dy, wd, rows_pad, kpad, ops.out_view_nchw(gx))
            elif flat_dgrad:
                src_full = run.vals[st.src]
                hp, wp_ = src_full.shape[1], src_full.shape[2]
                dfull = ops.alloc_flat_output(n, hp, wp_, dyp.shape[2], src_full.shape[3], dev)
                k = conv.kernel_size[0]
                if _toeplitz_last(conv, st.transposed, cs):
                    # few output channels (c7s1-3): dy is the 8-channel "image" of a Toeplitz convolution with the
                    # flipped filter producing the 64-channel gradient of the padded input
                    wr, rows_r = _pack_cache.get_toeplitz(conv.weight, False, True)
                    ops.conv2d_toeplitz_fwd(dyp, wr, rows_r, k, k, ops.out_view_nhwc(dfull, ci))
                elif cs <= 16 and k * cs <= 64 and conv.dilation[0] == 1:
                    # few output channels (c7s1-3): the zero-haloed dy carries 8 channels per pixel, so one K
                    # block covers a whole filter row (row-packed operand, taps reversed at packing time)
                    wr, rows_r, kpad_r = _pack_cache.get(conv.weight, st.transposed, cs, flipped=True)
                    ops.conv2d_fwd(ops.geom(k, k, 1, 0, 0, 1, False, cs), dyp, wr, rows_r, kpad_r,
                                   ops.out_view_nhwc(dfull, ci))
                else:
                    gflip = ops.geom(k, k, 1, 0, 0, conv.dilation[0], False, 0, True)
                    ops.conv2d_fwd(gflip, dyp, wd, rows_pad, kpad, ops.out_view_nhwc(dfull, ci))
                dpad[st.src] = dfull
                if DEBUG_RECORD is not None:
                    DEBUG_RECORD[('dfull', st.src)] = dfull.clone()
            else:
                src_full = run.vals[st.src]
                if materialised:
                    dfull = torch.empty_like(src_full)
                    ops.conv2d_fwd(gd, dy, wd, rows_pad, kpad, ops.out_view_nhwc(dfull, ci))
                else:
                    dfull = torch.empty_like(src_full)
                    hl = plan.halo[st.src]
                    if hl:
                        raise NotImplementedError("halo on a value read by a zero-padded convolution")
                    ops.conv2d_fwd(gd, dy, wd, rows_pad, kpad, ops.out_view_nhwc(dfull, ci))
                dpad[st.src] = dfull
                if DEBUG_RECORD is not None:
                    DEBUG_RECORD[('dfull', st.src)] = dfull.clone()
        # free saved tensors of this stage early
        run.y[idx] = None
    side.join()
    return gx, grads


class NetFunction(torch.autograd.Function):
    """One autograd node for a whole network call."""

    @staticmethod
    def forward(ctx, plan, training, x, *params):
        need_input_grad = bool(ctx.needs_input_grad[2])
        out, run = forward(plan, x.detach(), training, need_input_grad, {})
        ctx.plan = plan
        ctx.run = run
        return out

    @staticmethod
    def backward(ctx, gout):
        plan, run = ctx.plan, ctx.run
        if run is None:
            raise RuntimeError("cdb200: backward called twice on the same network call")
        needs = {p: bool(ctx.needs_input_grad[3 + i]) for i, p in enumerate(plan.params)}
        gx, grads = backward(plan, run, gout, ctx.needs_input_grad[2], needs)
        ctx.run = None
        return (None, None, gx) + tuple(grads.get(p) if needs[p] else None for p in plan.params)


def run_network(module, plan, x):
    """Executes `plan` for `module` on x with autograd support."""
    params = plan.params
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
        return NetFunction.apply(plan, module.training, x, *params)
    out, _ = forward(plan, x.detach(), module.training, False, {})
    return out
