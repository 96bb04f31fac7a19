"""K6 parity: fused LSGAN-MSE / BCE / L1 losses (value and gradient) against torch fp32."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("target", [1.0, 0.0])
def test_mse_const(target):
    from cycle_depth_estimation_b200 import losses
    x = torch.randn(8, 1, 30, 30, device="cuda", requires_grad=True)
    (losses.mse_const(x, target) * 3.0).backward()
    xr = x.detach().clone().requires_grad_(True)
    ref = F.mse_loss(xr, torch.full_like(xr, target)) * 3.0
    ref.backward()
    assert abs(float(losses.mse_const(x, target)) * 3.0 - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_l2(x.grad, xr.grad) < 1e-6


def test_bce_const():
    from cycle_depth_estimation_b200 import losses
    x = torch.rand(16, 1, 30, 30, device="cuda").clamp(1e-4, 1 - 1e-4).requires_grad_(True)
    losses.bce_const(x, 1.0).backward()
    xr = x.detach().clone().requires_grad_(True)
    ref = F.binary_cross_entropy(xr, torch.ones_like(xr))
    ref.backward()
    assert abs(float(losses.bce_const(x, 1.0)) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_l2(x.grad, xr.grad) < 1e-5


def test_l1_both_sides():
    from cycle_depth_estimation_b200 import losses
    a = torch.randn(8, 3, 64, 64, device="cuda", requires_grad=True)
    b = torch.randn(8, 3, 64, 64, device="cuda", requires_grad=True)
    (losses.l1(a, b) * 10.0).backward()
    ar, br = a.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    ref = F.l1_loss(ar, br) * 10.0
    ref.backward()
    assert abs(float(losses.l1(a, b)) * 10.0 - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_l2(a.grad, ar.grad) < 1e-6 and rel_l2(b.grad, br.grad) < 1e-6
