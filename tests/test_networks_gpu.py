"""Module-level parity on the GPU: the drop-in networks (fused B200 engine, bf16 tensor cores) against
the fp32 oracle restatement on identical weights and inputs.

Tolerances (relative L2):
* activations / outputs vs the pure fp32 oracle: <= 2e-2 (north_star, bf16 path);
* every backward KERNEL (wgrad, dgrad, norm/activation backward) on identical inputs vs torch fp32:
  tests/test_conv_gpu.py, tests/test_norm_gpu.py (<= 1e-3 .. 4e-3) — this is the parity bar of the kernels;
* network gradients vs the oracle evaluated with the SAME storage rounding (bf16 activations / weights,
  fp32 accumulation; straight-through gradient): same envelope as below, but consistently closer (the
  emulation removes part of the branch flips described next; values on a bf16 rounding boundary still
  round differently);
* the same network topology WITHOUT activation branches: <= 2e-2 end to end (pins the engine wiring);
* gradients vs the pure fp32 oracle: <= GRAD_FLIP_TOL. A reduced-precision forward perturbs
  pre-activations by ~0.5 %, which flips the ReLU / LeakyReLU branch of the ~0.4 % of elements that sit
  that close to zero; each flip changes that element's gradient by 80-100 %, i.e. ~5 % relative L2 per
  activation layer whatever the backward kernel does (measured: 5.0 / 6.6 / 7.8 / 8.4 % after 1..4
  layers of the PatchGAN). torch.autocast(bf16) on the reference shows the same. The bound documents
  that envelope; it is not the parity bar of the kernels.
Gradients of biases in front of an InstanceNorm are mathematically zero (SURVEY B-4) and are compared
against an absolute floor."""
import pytest
import torch

from helpers import TOL_BF16, leaf_state, quiet, rel_l2, seeded_image, true_fp32
from oracle import networks_oracle as O

pytestmark = pytest.mark.gpu

GRAD_FLIP_TOL = 0.25


def _build_G(n_blocks, ngf=64):
    from cycle_depth_estimation_b200 import networks as N
    torch.manual_seed(0)
    with quiet():
        net = N.ResnetGenerator(3, 3, ngf, norm_layer=N.get_norm_layer('instance'), n_blocks=n_blocks)
        N.init_weights(net, 'normal', 0.02)
    return net.cuda()


def _build_D(input_nc=3):
    from cycle_depth_estimation_b200 import networks as N
    torch.manual_seed(1)
    with quiet():
        net = N.define_D(input_nc, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cuda'])
    return net


def _compare_grads(net, sd, tol=TOL_BF16):
    worst = ("", 0.0)
    named = dict(net.named_parameters())
    for k, ref in sd.items():
        if not ref.requires_grad:
            continue
        got = named[k].grad
        assert got is not None, "missing gradient for " + k
        if ref.grad is None:
            continue
        if k.endswith(".bias") and float(got.abs().max()) == 0.0:
            # bias in front of a normalisation: cancelled exactly here, rounding noise in the reference
            wref = sd[k[:-4] + "weight"].grad
            assert float(ref.grad.double().norm()) <= 1e-4 * float(wref.double().norm()), k
            continue
        err = rel_l2(got, ref.grad)
        if err > worst[1]:
            worst = (k, err)
        assert err <= tol, (k, err)
    return worst


@pytest.mark.parametrize("n_blocks,size,batch", [(2, 64, 2), (9, 256, 1)])
def test_resnet_generator_forward(n_blocks, size, batch):
    net = _build_G(n_blocks)
    x = seeded_image(batch, 3, size, size)
    with torch.no_grad():
        got = net(x)
    with true_fp32(), torch.no_grad():
        ref = O.resnet_generator(net.state_dict(), x, n_blocks)
    assert got.shape == ref.shape and got.dtype == torch.float32
    err = rel_l2(got, ref)
    assert err <= TOL_BF16, err


def _backward_case(net, oracle_fn, x0, gout):
    x = x0.clone().requires_grad_(True)
    out = net(x)
    (out * gout).sum().backward()
    results = {}
    for mode in ("fp32", "bf16_storage"):
        sd = leaf_state(net)
        xr = x0.clone().requires_grad_(True)
        with true_fp32():
            if mode == "fp32":
                ref = oracle_fn(sd, xr)
            else:
                with O.emulate_bf16_storage():
                    ref = oracle_fn(sd, xr)
            (ref * gout).sum().backward()
        results[mode] = (ref.detach(), xr.grad, sd)
    ref, gx, sd = results["fp32"]
    assert rel_l2(out, ref) <= TOL_BF16, rel_l2(out, ref)
    assert rel_l2(x.grad, gx) <= GRAD_FLIP_TOL, rel_l2(x.grad, gx)
    _compare_grads(net, sd, GRAD_FLIP_TOL)
    ref, gx, sd = results["bf16_storage"]
    assert rel_l2(out, ref) <= TOL_BF16, rel_l2(out, ref)
    assert rel_l2(x.grad, gx) <= GRAD_FLIP_TOL, rel_l2(x.grad, gx)
    _compare_grads(net, sd, GRAD_FLIP_TOL)


def test_resnet_generator_backward():
    n_blocks = 3
    net = _build_G(n_blocks)
    _backward_case(net, lambda sd, x: O.resnet_generator(sd, x, n_blocks), seeded_image(2, 3, 64, 64),
                   seeded_image(2, 3, 64, 64, seed=7))


def test_nlayer_discriminator_forward_backward():
    net = _build_D()
    assert net(seeded_image(2, 3, 128, 128)).shape == (2, 1, 14, 14)
    _backward_case(net, lambda sd, x: O.nlayer_discriminator(sd, x), seeded_image(2, 3, 128, 128),
                   seeded_image(2, 1, 14, 14, seed=9))


def test_frozen_discriminator_gives_only_input_grad():
    net = _build_D()
    for p in net.parameters():
        p.requires_grad_(False)
    x = seeded_image(1, 3, 64, 64).requires_grad_(True)
    net(x).sum().backward()
    assert x.grad is not None and all(p.grad is None for p in net.parameters())


def test_state_dict_roundtrip_with_reference_layout():
    net = _build_G(9)
    keys = list(net.state_dict().keys())
    assert len(keys) == 48 and keys[0] == "model.1.weight" and keys[-1] == "model.26.bias"
    other = _build_G(9)
    other.load_state_dict({("module." + k)[7:]: v for k, v in net.state_dict().items()}, strict=True)


def test_no_cpu_path():
    from cycle_depth_estimation_b200 import networks as N
    with quiet():
        net = N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu'])
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 64, 64))


def _torch_run(mods, x):
    """Plain torch evaluation of a module list (stock leaf modules; residual blocks via conv_block)."""
    for m in mods:
        if hasattr(m, "conv_block"):
            x = x + _torch_run(list(m.conv_block.children()), x)
        else:
            x = m(x)
    return x


def test_engine_glue_without_activation_branches_is_tight():
    """Same generator topology with every ReLU removed (InstanceNorm keeps it non-trivial): no branch can
    flip, so the end-to-end gradients of the fused engine must meet the 2e-2 bar against fp32 torch.
    This pins the engine's wiring (halo folds, residual / skip gradients, stride-2 and transposed stages,
    first / last layer handling) independently of the ReLU-flip envelope."""
    import torch.nn as nn
    from cycle_depth_estimation_b200 import networks as N

    def strip(seq):
        out = []
        for m in seq.children():
            if isinstance(m, nn.ReLU):
                continue
            if hasattr(m, "conv_block"):
                m.conv_block = nn.Sequential(*strip(m.conv_block))
            out.append(m)
        return out

    net = _build_G(3)
    net.model = nn.Sequential(*strip(net.model))
    net.__dict__.pop('_cdb_plan', None)
    x0 = seeded_image(2, 3, 64, 64)
    gout = seeded_image(2, 3, 64, 64, seed=7)
    x = x0.clone().requires_grad_(True)
    out = net(x)
    (out * gout).sum().backward()
    got = {k: p.grad.clone() for k, p in net.named_parameters()}
    for p in net.parameters():
        p.grad = None
    xr = x0.clone().requires_grad_(True)
    with true_fp32():
        ref = _torch_run(list(net.model.children()), xr)
        (ref * gout).sum().backward()
    assert rel_l2(out, ref) <= TOL_BF16, rel_l2(out, ref)
    assert rel_l2(x.grad, xr.grad) <= TOL_BF16, rel_l2(x.grad, xr.grad)
    worst = ("", 0.0)
    for k, p in net.named_parameters():
        if k.endswith(".bias") and float(got[k].abs().max()) == 0.0:
            continue
        err = rel_l2(got[k], p.grad)
        if k.endswith(".bias"):
            # a bias gradient is a plain sum over 10^5 mixed-sign terms: cancellation makes its RELATIVE error the
            # most sensitive of all (observed 1.6e-2 .. 2.3e-2 run to run); allow 3e-2 there
            assert err <= 1.5 * TOL_BF16, (k, err)
            continue
        worst = max(worst, (k, err), key=lambda t: t[1])
    assert worst[1] <= TOL_BF16, worst
