"""CPU-side checks: the C-ABI library loads and exports every symbol include/cdb200.h declares, the
drop-in modules keep the reference's constructor / state_dict contract, and the product refuses to run
without CUDA (no silent fallback)."""
import argparse
import re

import pytest
import torch

from helpers import quiet


def test_library_exports_every_declared_symbol():
    from cycle_depth_estimation_b200 import _lib
    L = _lib.lib()
    syms = _lib.exported_symbols_from_header()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), s
    assert L.cdb_version() >= 100
    assert L.cdb_launch_count() == 0


def test_error_convention_without_gpu():
    """Bad descriptors are rejected on the host (negative status + message) before any CUDA call."""
    import ctypes as C
    from cycle_depth_estimation_b200 import _lib
    L = _lib.lib()
    rc = L.cdb_conv2d_fwd(None, None, None, 0, 0, None, None, None)
    assert rc == -1 and b"null" in L.cdb_last_error()
    g = _lib.CdbConvGeom(3, 3, 3, 0, 0, 1, 0, 0)  # stride 3 is unsupported
    x = _lib.CdbAct(16, 1, 8, 8, 8, 512, 64, 8, _lib.BF16, 0)
    y = _lib.CdbOut(16, 1, 2, 2, 8, 8, _lib.BF16, 32, 16, 8, 1)
    rc = L.cdb_conv2d_fwd(C.byref(g), C.byref(x), C.c_void_p(16), 16, 64, C.byref(y), None, None)
    assert rc == -2
    with pytest.raises(NotImplementedError):
        _lib.check(rc)
    x.sw = 4  # misaligned pixel stride
    g.stride = 1
    assert L.cdb_conv2d_fwd(C.byref(g), C.byref(x), C.c_void_p(16), 16, 64, C.byref(y), None, None) == -3


def test_ops_refuse_cpu_tensors():
    from cycle_depth_estimation_b200 import losses, networks as N
    with quiet():
        d = N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu'])
    with pytest.raises(RuntimeError):
        d(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        losses.l1(torch.zeros(4), torch.zeros(4))


def test_constructor_contract_and_state_dict_keys():
    from cycle_depth_estimation_b200 import networks as N
    with quiet():
        g = N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu'])
        d = N.define_D(3, 64, 'basic', 3, 'instance', False, 'normal', 0.02, ['cpu'])
        u = N.define_G(3, 3, 64, 'unet_256', 'batch', True, 'normal', 0.02, ['cpu'])
    assert len(g.state_dict()) == 48 and sum(p.numel() for p in g.parameters()) == 11378179
    assert list(d.state_dict())[:2] == ['model.0.weight', 'model.0.bias']
    assert sum(p.numel() for p in d.parameters()) == 2764737
    assert sum(p.numel() for p in u.parameters()) == 54413955
    assert any(k.endswith('num_batches_tracked') for k in u.state_dict())
    # DataParallel-prefixed checkpoints load after stripping the prefix, strictly
    g.load_state_dict({k: v for k, v in (("module." + k, v) for k, v in g.state_dict().items())
                       for k in [k[len("module."):]]}, strict=True)
    with pytest.raises(NotImplementedError):
        N.define_G(3, 3, 64, '3blocks', 'instance', False, 'normal', 0.02, ['cpu'])
    with pytest.raises(IndexError):
        with quiet():
            N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, [])
    sched = N.get_scheduler(torch.optim.SGD(d.parameters(), lr=1.0), argparse.Namespace(lr_policy='lambda'))
    assert sched.get_last_lr() == [1.0]
    assert isinstance(N.get_scheduler(None, argparse.Namespace(lr_policy='nope')), NotImplementedError)


def test_engine_plan_of_the_resnet_generator():
    from cycle_depth_estimation_b200 import engine, networks as N
    with quiet():
        g = N.define_G(3, 3, 64, 'resnet_9blocks', 'instance', False, 'normal', 0.02, ['cpu'])
    plan = g._plan()
    assert len(plan.stages) == 24
    assert [s.reflect for s in plan.stages[:4]] == [3, 0, 0, 1] and plan.stages[-1].reflect == 3
    assert [s.res is not None for s in plan.stages[3:21]] == [False, True] * 9
    assert [s.transposed for s in plan.stages[21:23]] == [True, True]
    assert plan.halo[0] == 3 and plan.halo[3] == 1 and plan.halo[21] == 0 and plan.halo[23] == 3
    assert len(plan.params) == 48


def test_gan_loss_buffers_in_state_dict():
    from cycle_depth_estimation_b200 import networks as N
    crit = N.GANLoss(use_lsgan=True)
    assert set(crit.state_dict()) == {'real_label', 'fake_label'}
    assert crit.get_target_tensor(torch.zeros(2, 1, 3, 3), True).shape == (2, 1, 3, 3)


def test_image_pool_plan_draws_the_same_decisions_as_query():
    """ImagePool.plan (the up-front decisions of the CUDA-graph step) consumes Python's ``random`` exactly like
    ImagePool.query: same trace, same RNG state afterwards."""
    import random
    import torch
    from cycle_depth_estimation_b200.image_pool import ImagePool
    random.seed(4321)
    a = ImagePool(5)
    for q in range(12):
        a.query(torch.zeros((3, 1, 2, 2)))
    state_a = random.getstate()
    random.seed(4321)
    b = ImagePool(5)
    plans = [b.plan(3) for _ in range(12)]
    assert a.trace == b.trace and random.getstate() == state_a
    flat = [p for plan in plans for p in plan]
    for (kind, idx), (ret, sto) in zip(a.trace, flat):
        assert (ret, sto) == {'fill': (-1, idx), 'swap': (idx, idx), 'pass': (-1, -1)}[kind]


def test_zero_arena_hands_out_disjoint_zero_views():
    import torch
    from cycle_depth_estimation_b200 import ops
    arena = ops.ZeroArena(torch.device('cpu'))
    a, b = arena.take((2, 3, 2)), arena.take((1, 5))
    big = arena.take((ops.ZeroArena.CHUNK + 8,))
    a += 1
    assert float(b.sum()) == 0.0 and float(big.sum()) == 0.0 and a.shape == (2, 3, 2) and big.numel() == ops.ZeroArena.CHUNK + 8


def test_unet_levels_and_networks5_module_trees():
    from cycle_depth_estimation_b200 import networks as N, networks5_ds as N5
    net = N.UnetGenerator(3, 3, 8, 64, N.get_norm_layer('batch'), True)
    lv = net._levels()
    assert len(lv) == 8 and lv[0][1] is None and lv[-1][1] is None and lv[-1][3] is not None
    assert [l[4] is not None for l in lv] == [False, False, False, False, True, True, True, False]
    d = N5._Discriminator(input_nc=128)
    assert d.model[1] is d.model[10] and 'model.10.weight' in d.state_dict()
    assert sum(p.numel() for p in N5.G_1().parameters()) == 615808
    assert sum(p.numel() for p in N5.General_net().parameters()) == 22808448
    assert sum(p.numel() for p in N5.R_dep().parameters()) == 52765086
    with __import__('pytest').raises(RuntimeError):
        N5.G_1()(__import__('torch').zeros(1, 3, 32, 32))     # CPU tensors are refused: no fallback path


def test_error_convention_of_the_entry_points_added_for_the_next_rows():
    """Null arguments / bad sizes are refused on the host side before anything touches a device (SURVEY 8(b) error
    convention) for the TF32, f3 and f4 entry points."""
    import ctypes as C
    from cycle_depth_estimation_b200 import _lib
    L = _lib.lib()
    assert L.cdb_pack_conv_weight_tf32(None, 8, 8, 3, 3, 1, None, None) == -1 and b"null" in L.cdb_last_error()
    assert L.cdb_round_tf32(None, C.c_int64(16), None) == -1
    assert L.cdb_round_tf32(None, C.c_int64(0), None) == 0
    assert L.cdb_scale(None, C.c_float(0.5), None, None) == -1
    assert L.cdb_nearest2x_fwd(None, None, None) == -1 and L.cdb_nearest2x_bwd(None, None, None) == -1
    assert L.cdb_tanh_fwd(None, None, None) == -1 and L.cdb_tanh_bwd(None, None, None, None) == -1
    assert L.cdb_depth_labels_workspace(3) == 24 and L.cdb_depth_labels_workspace(0) == 0
    assert L.cdb_depth_labels(None, 1, C.c_int64(16), None, None, None, C.c_size_t(0), None) == -1
    assert L.cdb_label_lut_i64(None, C.c_int64(4), None, None, None) == -1
    assert L.cdb_image_normalize_u8(None, 1, C.c_int64(4), 3, C.c_float(0.5), C.c_float(0.5), None, None) == -1
    x = _lib.CdbAct(16, 1, 4, 4, 8, 128, 32, 8, _lib.BF16, 0)
    y = _lib.CdbAct(16, 1, 4, 4, 16, 256, 64, 16, _lib.BF16, 0)
    assert L.cdb_scale(C.byref(x), C.c_float(0.5), C.byref(y), None) == -1 and b"shapes" in L.cdb_last_error()
    assert L.cdb_nearest2x_fwd(C.byref(x), C.byref(x), None) == -1
    f = _lib.CdbAct(16, 1, 4, 4, 8, 128, 32, 8, 7, 0)
    assert L.cdb_tanh_fwd(C.byref(f), C.byref(f), None) == -2          # unknown storage type: unsupported, not a fallback
    f32 = _lib.CdbAct(16, 1, 4, 4, 8, 128, 32, 8, _lib.F32, 0)
    assert L.cdb_split_tf32(C.byref(x), C.byref(f32), 0, None) == -2   # the TF32 operand split reads fp32 views only
    assert L.cdb_split_tf32(C.byref(f32), C.byref(f32), 0, None) == -1 and b"3c" in L.cdb_last_error()
    assert L.cdb_split_tf32(C.byref(f32), C.byref(f32), 9, None) == -1
    assert L.cdb_dropout_dev(C.byref(x), C.byref(x), C.c_uint64(1), None, C.c_float(0.5), None) == -1
    assert L.cdb_adam_pack_multi(None, 0, C.c_float(1e-3), None, C.c_float(0.5), C.c_float(0.999), C.c_float(1e-8), 1,
                                 None, None) == -1


def test_fused_adam_is_a_torch_optimizer_that_schedulers_accept():
    """networks.get_scheduler (models/networks.py:24-38) and update_learning_rate (models/base_model.py:91-95) act on
    FusedAdam like on torch.optim.Adam; its state_dict uses torch.optim.Adam's keys and round-trips."""
    import argparse
    import torch
    from cycle_depth_estimation_b200 import networks as N
    from cycle_depth_estimation_b200.cycle_gan_model import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4, 3))
    opt = FusedAdam([p], lr=2e-4, betas=(0.5, 0.999))
    assert isinstance(opt, torch.optim.Optimizer)
    assert opt.param_groups[0]['lr'] == 2e-4 and opt.param_groups[0]['betas'] == (0.5, 0.999)
    for policy in ('lambda', 'step', 'plateau', 'cosine'):
        o = FusedAdam([p], lr=2e-4, betas=(0.5, 0.999))
        sch = N.get_scheduler(o, argparse.Namespace(lr_policy=policy, lr_decay_iters=2, niter=10))
        assert not isinstance(sch, Exception), policy
    sch = N.get_scheduler(opt, argparse.Namespace(lr_policy='lambda'))
    lrs = []
    for _ in range(14):
        sch.step()
        lrs.append(opt.param_groups[0]['lr'])
    assert lrs[9] == 2e-4 and abs(lrs[12] - 2e-4 * (1 - 3 / 30.0)) < 1e-12     # constant for 10 epochs, then linear decay
    sd = opt.state_dict()
    assert set(sd.keys()) == {'state', 'param_groups'}
    other = FusedAdam([torch.nn.Parameter(torch.zeros(4, 3))], lr=1.0)
    other.load_state_dict(sd)
    assert other.param_groups[0]['lr'] == opt.param_groups[0]['lr']
    opt.zero_grad()
    assert p.grad is None


def test_image_pool_trace_is_bounded_and_plan_handles_an_empty_pool():
    from cycle_depth_estimation_b200.image_pool import ImagePool, _Trace
    t = _Trace()
    for i in range(_Trace.MAX + 10):
        t.append(('pass', -1))
    assert len(t) <= _Trace.MAX and isinstance(t, list)
    assert ImagePool(0).plan(3) == [(-1, -1)] * 3


def test_precision_switch():
    from cycle_depth_estimation_b200 import ops
    assert ops.get_precision() == 'bf16'
    with ops.precision('tf32x3'):
        assert ops.get_precision() == 'tf32x3'
        with ops.precision('tf32'):
            assert ops.get_precision() == 'tf32'
        assert ops.get_precision() == 'tf32x3'
    assert ops.get_precision() == 'bf16'
    import pytest
    with pytest.raises(ValueError):
        ops.set_precision('fp64')
