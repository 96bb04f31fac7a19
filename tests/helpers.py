"""Shared helpers of the test-suite."""
import contextlib
import io

import torch

TOL_BF16 = 2e-2   # north_star: relative L2 <= 2e-2 for the bf16 path (activations, losses, gradients)


def rel_l2(a, b, floor=0.0):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / max(float(b.norm()), floor, 1e-30))


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


@contextlib.contextmanager
def true_fp32():
    """The oracle must be evaluated in real fp32 on the GPU (no TF32 inside cuDNN / cuBLAS)."""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def seeded_image(n, c, h, w, seed=1234, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand((n, c, h, w), generator=g) * 2 - 1).to(device)


def leaf_state(module):
    """state_dict of leaf tensors requiring grad, keyed like the reference, sharing no storage."""
    return {k: v.detach().clone().requires_grad_(v.is_floating_point() and 'running_' not in k)
            for k, v in module.state_dict().items()}
