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
    """torch.nn.L1Loss()(a, b)."""
    return _L1.apply(a, b)


class L1Loss(torch.nn.Module):
    """Drop-in for torch.nn.L1Loss() at the reference's call sites."""

    def forward(self, input, target):
        return l1(input, target)
