"""Fused loss functions (K6): forward scalar and gradient in one kernel, autograd-aware.

Replace nn.MSELoss / nn.BCELoss against a constant label inside GANLoss (models/networks.py:119-138)
and nn.L1Loss of the cycle / identity terms (models/cycle_gan_model.py:63-64,119-134).
"""
import torch

from . import ops


class _ConstTargetLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, kind):
        xc = x.detach().contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        (ops.loss_mse_const if kind == 'mse' else ops.loss_bce_const)(xc, float(target), 1.0, loss, grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return ops.scale_by_scalar(ctx.grad, g.contiguous().float()), None, None


class _L1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ac, bc = a.detach().contiguous(), b.detach().contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=a.device)
        grad = torch.empty_like(ac) if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) else None
        ops.loss_l1(ac, bc, 1.0, loss, grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        ga = ops.scale_by_scalar(ctx.grad, g.contiguous().float())
        gb = None
        if ctx.needs_input_grad[1]:
            gb = ops.scale_by_scalar(ctx.grad, (-g).contiguous().float())
        return (ga if ctx.needs_input_grad[0] else None), gb


def mse_const(x, target):
    """mean((x - target)^2) for a python-float target (LSGAN)."""
    return _ConstTargetLoss.apply(x, target, 'mse')


def bce_const(x, target):
    """torch.nn.BCELoss(x, full_like(x, target)) (vanilla GAN on sigmoid outputs)."""
    return _ConstTargetLoss.apply(x, target, 'bce')


def l1(a, b):
    """torch.nn.L1Loss()(a, b), including its broadcasting of mismatched shapes (the seg/depth step compares a
    [B,1,H,W] prediction with a [B,H,W] label, new_multi/model5.py:532,573, which broadcasts to [B,B,H,W])."""
    if a.shape != b.shape:
        a, b = torch.broadcast_tensors(a, b)
    return _L1.apply(a, b)


class _CE2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        lc = logits.detach().contiguous()
        acc = torch.zeros((2,), dtype=torch.float32, device=logits.device)
        grad = torch.empty_like(lc) if ctx.needs_input_grad[0] else None
        ops.loss_ce2d(lc, labels.contiguous(), int(ignore_index), acc, grad)
        ctx.grad, ctx.count = grad, acc[1:2]
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, g):
        return ops.scale_by_scalar(ctx.grad, (g.reshape(1).float() / ctx.count).contiguous()), None, None


def cross_entropy2d(logits, labels, ignore_index=255):
    """torch.nn.CrossEntropyLoss(ignore_index=...)(logits [N,C,H,W], labels [N,H,W]) (new_multi/model5.py:281):
    mean over the non-ignored pixels; value and gradient from one fused kernel."""
    if labels.dtype != torch.int64:
        raise TypeError("int64 class labels expected")
    return _CE2d.apply(logits, labels, ignore_index)


class CrossEntropyLoss(torch.nn.Module):
    """Drop-in for torch.nn.CrossEntropyLoss(size_average=True, ignore_index=255) at new_multi/model5.py:281."""

    def __init__(self, size_average=True, ignore_index=255):
        super(CrossEntropyLoss, self).__init__()
        if not size_average:
            raise NotImplementedError("sum reduction")
        self.ignore_index = ignore_index

    def forward(self, input, target):
        return cross_entropy2d(input, target, self.ignore_index)


class _BCEDep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        xc, tc = x.detach().contiguous(), target.detach().contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        ops.loss_bcedep(xc, tc, 50.0, loss, grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return ops.scale_by_scalar(ctx.grad, g.contiguous().float()), None


def depth_bin_masks(target):
    """get_masks (new_multi/networks5_ds.py:973-982): exact-equality masks of the +1 / -1 depth bins."""
    t = target.detach()
    return (t == 1).to(t.dtype), (t == -1).to(t.dtype)


def bcedep(input, target, o_m=None, z_m=None):
    """BCEDepLoss (new_multi/networks5_ds.py:947-956) for masks produced by get_masks(target) — the kernel
    recomputes them from the target, so o_m / z_m are accepted for signature compatibility only."""
    if input.dim() != 4 or target.dim() != 4 or input.shape[1] != 1:
        raise NotImplementedError("BCEDepLoss: input [B,1,H,W] against target [B,K,H,W] expected")
    return _BCEDep.apply(input, target)


class L1Loss(torch.nn.Module):
    """Drop-in for torch.nn.L1Loss() at the reference's call sites."""

    def forward(self, input, target):
        return l1(input, target)
