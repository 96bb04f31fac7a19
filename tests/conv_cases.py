"""Convolution parity cases shared by tests/test_conv_gpu.py and tools/gpu_probe.py.

Each case builds seeded inputs, runs the tcgen05 path through the C ABI and compares with a plain
PyTorch fp32 evaluation of the same convolution on the SAME bf16-rounded operands, so the only
differences are accumulation order and the rounding of the stored result.
Tolerances (relative L2): bf16 outputs 4e-3 (one rounding, 2^-9), fp32 outputs 2e-5.
"""
import torch
import torch.nn.functional as F

from cycle_depth_estimation_b200 import ops

TOL_BF16 = 4e-3
TOL_F32 = 2e-5


def rel_l2(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _rand(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


def nhwc(x_nchw, cstore=None):
    """NCHW (any float) -> contiguous NHWC bf16 with channels zero-padded to cstore."""
    n, c, h, w = x_nchw.shape
    cs = cstore or ops.round_up(c, 8)
    out = torch.zeros((n, h, w, cs), dtype=torch.bfloat16, device=x_nchw.device)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1)
    return out


def conv_fwd_case(n, cin, cout, h, w, k, stride=1, pad=0, dil=1, transposed=False, out_pad=0,
                  out_mode="nhwc", bias=False, act=ops.ACT_NONE, stats=False, seed=0, flip=False, pitched=False):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    x = _rand((n, cin, h, w), gen)
    if transposed:
        w4 = _rand((cin, cout, k, k), gen, 0.05)
        ref = F.conv_transpose2d(x.float(), w4.float(), None, stride, pad, out_pad, 1, dil)
    else:
        w4 = _rand((cout, cin, k, k), gen, 0.05)
        ref = F.conv2d(x.float(), w4.float().flip(2, 3) if flip else w4.float(), None, stride, pad, dil)
    b = None
    if bias:
        b = torch.randn(cout, generator=gen, device="cuda")
        ref = ref + b.view(1, -1, 1, 1)
    if act == ops.ACT_RELU:
        ref = torch.relu(ref)
    elif act == ops.ACT_LEAKY:
        ref = F.leaky_relu(ref, 0.2)
    elif act == ops.ACT_TANH:
        ref = torch.tanh(ref)
    p, q = ref.shape[2], ref.shape[3]
    xs = nhwc(x)
    wp, rows_pad, kpad = ops.pack_conv_weight(w4.float().contiguous(), rows_are_dim0=not transposed)
    g = ops.geom(k, k, stride, pad, pad, dil, transposed, 0, flip)
    st = torch.zeros((n, cout, 2), dtype=torch.float32, device="cuda") if stats else None
    if out_mode == "nhwc":
        cs = ops.round_up(cout, 8)
        if pitched:
            y = ops.alloc_flat_output(n, p, q, w, cs, "cuda")
            y.fill_(float("nan"))
        else:
            y = torch.full((n, p, q, cs), float("nan"), dtype=torch.bfloat16, device="cuda")
        ops.conv2d_fwd(g, xs, wp, rows_pad, kpad, ops.out_view_nhwc(y, cout), b, act, 0.2, st)
        got = y[..., :cout].permute(0, 3, 1, 2).float()
        tol = TOL_BF16
        pad_ok = bool((y[..., cout:] == 0).all()) if cs > cout else True
    else:
        y = torch.full((n, cout, p, q), float("nan"), dtype=torch.float32, device="cuda")
        ops.conv2d_fwd(g, xs, wp, rows_pad, kpad, ops.out_view_nchw(y), b, act, 0.2, st)
        got = y
        tol = TOL_F32
        pad_ok = True
    torch.cuda.synchronize()
    err = rel_l2(got, ref)
    res = {"err": err, "tol": tol, "ok": err <= tol and pad_ok and bool(torch.isfinite(got).all())}
    if stats:
        s_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1)
        res["stats_err"] = rel_l2(st, s_ref)
        res["ok"] = res["ok"] and res["stats_err"] < 1e-3
    return res


TOL_TF32 = 1e-3  # north-star gate for the TF32 variant (relative L2 against the fp32 reference)


def nhwc_f32(x_nchw, cstore=None):
    n, c, h, w = x_nchw.shape
    cs = cstore or ops.round_up(c, 4)
    out = torch.zeros((n, h, w, cs), dtype=torch.float32, device=x_nchw.device)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1)
    return out


def conv_fwd_tf32_case(n, cin, cout, h, w, k, stride=1, pad=0, dil=1, transposed=False, out_pad=0, out_mode="nhwc",
                       bias=False, act=ops.ACT_NONE, stats=False, seed=0, flip=False, round_x=True, round_out=False):
    """TF32 variant of K1/K2: full-precision fp32 operands in, reference evaluated in float64 on the UNROUNDED
    operands, so the error is the whole TF32 error (operand rounding + fp32 accumulation)."""
    gen = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((n, cin, h, w), generator=gen, device="cuda")
    if transposed:
        w4 = torch.randn((cin, cout, k, k), generator=gen, device="cuda") * 0.05
        ref = F.conv_transpose2d(x.double(), w4.double(), None, stride, pad, out_pad, 1, dil)
    else:
        w4 = torch.randn((cout, cin, k, k), generator=gen, device="cuda") * 0.05
        ref = F.conv2d(x.double(), w4.double().flip(2, 3) if flip else w4.double(), None, stride, pad, dil)
    b = None
    if bias:
        b = torch.randn(cout, generator=gen, device="cuda")
        ref = ref + b.double().view(1, -1, 1, 1)
    if act == ops.ACT_LEAKY:
        ref = F.leaky_relu(ref, 0.2)
    elif act == ops.ACT_TANH:
        ref = torch.tanh(ref)
    p, q = ref.shape[2], ref.shape[3]
    xs = nhwc_f32(x)
    if round_x:
        ops.round_tf32_(xs)
    wp, rows_pad, kpad = ops.pack_conv_weight_tf32(w4.contiguous(), rows_are_dim0=not transposed)
    g = ops.geom(k, k, stride, pad, pad, dil, transposed, 0, flip)
    st = torch.zeros((n, cout, 2), dtype=torch.float32, device="cuda") if stats else None
    flags = ops.EP_ROUND_TF32 if round_out else 0
    if out_mode == "nhwc":
        cs = ops.round_up(cout, 4)
        y = torch.full((n, p, q, cs), float("nan"), dtype=torch.float32, device="cuda")
        ops.conv2d_fwd(g, xs, wp, rows_pad, kpad, ops.out_view_nhwc(y, cout), b, act, 0.2, st, flags)
        got = y[..., :cout].permute(0, 3, 1, 2)
        pad_ok = bool((y[..., cout:] == 0).all()) if cs > cout else True
    else:
        y = torch.full((n, cout, p, q), float("nan"), dtype=torch.float32, device="cuda")
        ops.conv2d_fwd(g, xs, wp, rows_pad, kpad, ops.out_view_nchw(y), b, act, 0.2, st, flags)
        got = y
        pad_ok = True
    torch.cuda.synchronize()
    err = rel_l2(got, ref)
    res = {"err": err, "tol": TOL_TF32, "ok": err <= TOL_TF32 and pad_ok and bool(torch.isfinite(got).all())}
    if round_out:  # every stored value must already be a TF32 number (low 13 mantissa bits clear)
        res["rounded"] = bool(((got.contiguous().view(torch.int32) & 0x1FFF) == 0).all())
        res["ok"] = res["ok"] and res["rounded"]
    if stats:
        s_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1)
        res["stats_err"] = rel_l2(st, s_ref)
        res["ok"] = res["ok"] and res["stats_err"] < 1e-3
    return res


def conv_wgrad_tf32_case(n, cin, cout, h, w, k, stride=1, pad=0, dil=1, transposed=False, out_pad=0,
                         accumulate=False, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((n, cin, h, w), generator=gen, device="cuda")
    xd = x.double()
    if transposed:
        w4 = torch.zeros((cin, cout, k, k), device="cuda", dtype=torch.float64, requires_grad=True)
        y = F.conv_transpose2d(xd, w4, None, stride, pad, out_pad, 1, dil)
    else:
        w4 = torch.zeros((cout, cin, k, k), device="cuda", dtype=torch.float64, requires_grad=True)
        y = F.conv2d(xd, w4, None, stride, pad, dil)
    dy = torch.randn(tuple(y.shape), generator=gen, device="cuda") * 0.1
    y.backward(dy.double())
    ref = w4.grad
    xs, dys = ops.round_tf32_(nhwc_f32(x)), ops.round_tf32_(nhwc_f32(dy))
    g = ops.geom(k, k, stride, pad, pad, dil, transposed, 0)
    base = (torch.randn(ref.shape, generator=gen, device="cuda") if accumulate
            else torch.zeros(ref.shape, device="cuda"))
    dw = base.clone()
    ops.conv2d_wgrad(g, xs, dys, dw, accumulate)
    torch.cuda.synchronize()
    err = rel_l2(dw - base if accumulate else dw, ref)
    return {"err": err, "tol": TOL_TF32, "ok": err <= TOL_TF32}


TF32_FWD_CASES = {
    "tf32_gemm_1x1_c64_n16": dict(n=1, cin=64, cout=16, h=8, w=16, k=1),
    "tf32_gemm_1x1_c128_n256": dict(n=2, cin=128, cout=256, h=16, w=32, k=1),
    "tf32_gemm_1x1_c96_n512_truncated_x": dict(n=1, cin=96, cout=512, h=16, w=16, k=1, round_x=False),
    "tf32_r256_zero_pad_stats": dict(n=2, cin=256, cout=256, h=64, w=64, k=3, pad=1, stats=True, round_out=True),
    "tf32_r256_prepadded_nchw": dict(n=1, cin=256, cout=256, h=66, w=66, k=3, out_mode="nchw"),
    "tf32_flat_flip_c128_bias_leaky": dict(n=2, cin=128, cout=128, h=40, w=50, k=3, flip=True, bias=True, act=ops.ACT_LEAKY, round_out=True),
    "tf32_flat_4x4_c256_512_stats": dict(n=2, cin=256, cout=512, h=34, w=34, k=4, stats=True),
    "tf32_conv3x3_s2_64_128": dict(n=2, cin=64, cout=128, h=32, w=32, k=3, stride=2, pad=1),
    "tf32_conv4x4_s2_patchgan": dict(n=2, cin=64, cout=128, h=64, w=64, k=4, stride=2, pad=1, bias=True, act=ops.ACT_LEAKY),
    "tf32_conv4x4_s1_cout1": dict(n=2, cin=512, cout=1, h=31, w=31, k=4, pad=1, bias=True, out_mode="nchw"),
    "tf32_conv3x3_dil2": dict(n=1, cin=64, cout=64, h=12, w=40, k=3, pad=2, dil=2),
    "tf32_conv7x7_cout3_tanh": dict(n=1, cin=64, cout=3, h=70, w=70, k=7, bias=True, act=ops.ACT_TANH, out_mode="nchw"),
    "tf32_conv_cin3_image": dict(n=1, cin=3, cout=64, h=38, w=38, k=7),
    "tf32_convT3x3_s2": dict(n=2, cin=256, cout=128, h=16, w=16, k=3, stride=2, pad=1, transposed=True, out_pad=1),
    "tf32_convT4x4_s2": dict(n=2, cin=128, cout=64, h=8, w=8, k=4, stride=2, pad=1, transposed=True),
    "tf32_dgrad_flip_3x3_pad2": dict(n=1, cin=256, cout=256, h=64, w=64, k=3, pad=2, flip=True),
}

TF32_WGRAD_CASES = {
    "tf32_wgrad_3x3_256": dict(n=2, cin=256, cout=256, h=34, w=34, k=3),
    "tf32_wgrad_3x3_pad1_128_64": dict(n=2, cin=128, cout=64, h=32, w=32, k=3, pad=1, accumulate=True),
    "tf32_wgrad_4x4_s2": dict(n=2, cin=64, cout=128, h=32, w=32, k=4, stride=2, pad=1),
    "tf32_wgrad_4x4_s1_cout1": dict(n=2, cin=512, cout=1, h=31, w=31, k=4, pad=1),
    "tf32_wgrad_convT3x3_s2": dict(n=2, cin=256, cout=128, h=16, w=16, k=3, stride=2, pad=1, transposed=True, out_pad=1),
    "tf32_wgrad_cin3_7x7": dict(n=2, cin=3, cout=64, h=38, w=38, k=7),
}


def conv_rowpack_case(n, cin, cout, h, w, k, stride, rowpack, seed=0):
    """Image-input layers: x holds `rowpack` channels per pixel, padding materialised by the caller."""
    gen = torch.Generator(device="cuda").manual_seed(seed)
    x = _rand((n, cin, h, w), gen)  # already padded image
    w4 = _rand((cout, cin, k, k), gen, 0.05)
    ref = F.conv2d(x.float(), w4.float(), None, stride, 0, 1)
    p, q = ref.shape[2], ref.shape[3]
    span = 64 // rowpack
    wneed = max(w, stride * (q - 1) + span)
    xs = torch.zeros((n, h, wneed, rowpack), dtype=torch.bfloat16, device="cuda")
    xs[:, :, :w, :cin] = x.permute(0, 2, 3, 1)
    wp, rows_pad, kpad = ops.pack_conv_weight(w4.float().contiguous(), True, rowpack)
    g = ops.geom(k, k, stride, 0, 0, 1, False, rowpack)
    cs = ops.round_up(cout, 8)
    y = torch.full((n, p, q, cs), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.conv2d_fwd(g, xs, wp, rows_pad, kpad, ops.out_view_nhwc(y, cout))
    torch.cuda.synchronize()
    got = y[..., :cout].permute(0, 3, 1, 2).float()
    err = rel_l2(got, ref)
    return {"err": err, "tol": TOL_BF16, "ok": err <= TOL_BF16}


def conv_wgrad_case(n, cin, cout, h, w, k, stride=1, pad=0, dil=1, transposed=False, out_pad=0,
                    rowpack=0, accumulate=False, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    x = _rand((n, cin, h, w), gen)
    xf = x.float().requires_grad_(False)
    if transposed:
        w4 = torch.zeros((cin, cout, k, k), device="cuda", requires_grad=True)
        y = F.conv_transpose2d(xf, w4, None, stride, pad, out_pad, 1, dil)
    else:
        w4 = torch.zeros((cout, cin, k, k), device="cuda", requires_grad=True)
        y = F.conv2d(xf, w4, None, stride, pad, dil)
    dy = _rand(tuple(y.shape), gen, 0.1)
    y.backward(dy.float())
    ref = w4.grad
    if rowpack:
        span = 64 // rowpack
        q = y.shape[3]
        wneed = max(w, stride * (q - 1) + span)
        xs = torch.zeros((n, h, wneed, rowpack), dtype=torch.bfloat16, device="cuda")
        xs[:, :, :w, :cin] = x.permute(0, 2, 3, 1)
    else:
        xs = nhwc(x)
    dys = nhwc(dy)
    g = ops.geom(k, k, stride, pad, pad, dil, transposed, rowpack)
    base = torch.randn(ref.shape, generator=gen, device="cuda") if accumulate else torch.zeros_like(ref)
    dw = base.clone()
    ops.conv2d_wgrad(g, xs, dys, dw, accumulate)
    torch.cuda.synchronize()
    err = rel_l2(dw - base if accumulate else dw, ref)
    return {"err": err, "tol": 1e-3, "ok": err <= 1e-3}


FWD_CASES = {
    # name: kwargs
    "gemm_1x1_c64_n16": dict(n=1, cin=64, cout=16, h=8, w=16, k=1),
    "gemm_1x1_c128_n256": dict(n=2, cin=128, cout=256, h=16, w=32, k=1),
    "gemm_1x1_c64_n512": dict(n=1, cin=64, cout=512, h=16, w=16, k=1),
    "conv3x3_zero_pad_256": dict(n=2, cin=256, cout=256, h=64, w=64, k=3, pad=1),
    "conv3x3_prepadded_stats": dict(n=2, cin=128, cout=128, h=34, w=34, k=3, stats=True),
    "conv3x3_s2_64_128": dict(n=2, cin=64, cout=128, h=32, w=32, k=3, stride=2, pad=1),
    "conv4x4_s2_patchgan": dict(n=2, cin=64, cout=128, h=64, w=64, k=4, stride=2, pad=1, bias=True, act=ops.ACT_LEAKY),
    "conv4x4_s1_31": dict(n=2, cin=256, cout=512, h=32, w=32, k=4, pad=1),
    "conv4x4_s1_cout1": dict(n=2, cin=512, cout=1, h=31, w=31, k=4, pad=1, bias=True, out_mode="nchw"),
    "conv3x3_dil2": dict(n=1, cin=64, cout=64, h=12, w=40, k=3, pad=2, dil=2),
    "conv7x7_cout3_tanh_nchw": dict(n=1, cin=64, cout=3, h=70, w=70, k=7, bias=True, act=ops.ACT_TANH, out_mode="nchw"),
    "conv_cin8": dict(n=1, cin=8, cout=64, h=16, w=16, k=3, pad=1),
    "conv_small_spatial_batchtile": dict(n=16, cin=512, cout=512, h=2, w=2, k=4, stride=2, pad=1),
    "convT3x3_s2": dict(n=2, cin=256, cout=128, h=16, w=16, k=3, stride=2, pad=1, transposed=True, out_pad=1),
    "convT4x4_s2": dict(n=2, cin=128, cout=64, h=8, w=8, k=4, stride=2, pad=1, transposed=True),
    "convT4x4_s2_from1x1": dict(n=4, cin=512, cout=512, h=1, w=1, k=4, stride=2, pad=1, transposed=True),
    "dgrad_like_3x3_pad2": dict(n=1, cin=256, cout=256, h=64, w=64, k=3, pad=2),
    # flat kernel (stride 1, materialised padding, contiguous input)
    "flat_r256_stats": dict(n=3, cin=256, cout=256, h=66, w=66, k=3, stats=True),
    "flat_r256_flip": dict(n=2, cin=256, cout=256, h=68, w=68, k=3, flip=True),
    "flat_3x3_c64_n128_dil2": dict(n=2, cin=64, cout=128, h=44, w=52, k=3, dil=2),
    "flat_7x7_c64_cout3_wide": dict(n=2, cin=64, cout=3, h=134, w=262, k=7, bias=True, act=ops.ACT_TANH, out_mode="nchw"),
    "flat_4x4_c512_cout1": dict(n=2, cin=512, cout=1, h=33, w=33, k=4, bias=True, out_mode="nchw"),
    "flat_1x1_c128_n320": dict(n=1, cin=128, cout=320, h=24, w=80, k=1),
    "generic_flip_s1_pad1": dict(n=1, cin=64, cout=64, h=20, w=20, k=3, pad=1, flip=True),
    # flat kernel, TMA-store output path (pitched output buffer)
    "flatfast_r256_stats": dict(n=3, cin=256, cout=256, h=66, w=66, k=3, stats=True, pitched=True),
    "flatfast_r256_flip": dict(n=2, cin=256, cout=256, h=68, w=68, k=3, flip=True, pitched=True),
    "flatfast_c128_bias_leaky": dict(n=2, cin=128, cout=128, h=40, w=50, k=3, bias=True, act=ops.ACT_LEAKY, pitched=True),
    "flatfast_cout72_stats": dict(n=2, cin=64, cout=72, h=30, w=34, k=3, stats=True, pitched=True),
    "flatfast_cout320": dict(n=1, cin=128, cout=320, h=26, w=82, k=3, pitched=True),
    "flatfast_4x4_c256_512": dict(n=2, cin=256, cout=512, h=34, w=34, k=4, stats=True, pitched=True),
}

ROWPACK_CASES = {
    "rowpack8_7x7": dict(n=2, cin=3, cout=64, h=70, w=70, k=7, stride=1, rowpack=8),
    "rowpack16_4x4_s2": dict(n=2, cin=3, cout=64, h=66, w=66, k=4, stride=2, rowpack=16),
    "rowpack16_4x4_s2_cin6": dict(n=1, cin=6, cout=64, h=34, w=34, k=4, stride=2, rowpack=16),
}

def conv_wgrad_fewcout_case(n, cin, cout, h, w, k, seed=0):
    """Weight gradient of a stride-1 convolution with <= 8 output channels through the transposed +
    row-packed form (x: padded input, dy: zero-haloed by k-1), as the engine does for the c7s1-3 layer."""
    gen = torch.Generator(device="cuda").manual_seed(seed)
    x = _rand((n, cin, h, w), gen)
    w4 = torch.zeros((cout, cin, k, k), device="cuda", requires_grad=True)
    y = F.conv2d(x.float(), w4)
    dy = _rand(tuple(y.shape), gen, 0.1)
    y.backward(dy.float())
    ho, wo = y.shape[2], y.shape[3]
    hz = k - 1
    dyp = torch.zeros((n, ho + 2 * hz, wo + 2 * hz + 8, 8), dtype=torch.bfloat16, device="cuda")
    dyp[:, hz:hz + ho, hz:hz + wo, :cout] = dy.permute(0, 2, 3, 1)
    tmp = torch.empty((cin, cout, k, k), dtype=torch.float32, device="cuda")
    ops.conv2d_wgrad(ops.geom(k, k, 1, 0, 0, 1, True, 8), nhwc(x), dyp, tmp, False)
    torch.cuda.synchronize()
    got = tmp.flip(2, 3).permute(1, 0, 2, 3)
    err = rel_l2(got, w4.grad)
    return {"err": err, "tol": 1e-3, "ok": err <= 1e-3}


WGRAD_FEWCOUT_CASES = {
    "wgrad_fewcout_7x7_64_3": dict(n=2, cin=64, cout=3, h=70, w=70, k=7),
    "wgrad_fewcout_3x3_128_1": dict(n=1, cin=128, cout=1, h=20, w=36, k=3),
}


WGRAD_CASES = {
    "wgrad_3x3_256": dict(n=2, cin=256, cout=256, h=34, w=34, k=3),
    "wgrad_3x3_pad1_128_64": dict(n=2, cin=128, cout=64, h=32, w=32, k=3, pad=1, accumulate=True),
    "wgrad_3x3_s2": dict(n=2, cin=64, cout=128, h=32, w=32, k=3, stride=2, pad=1),
    "wgrad_4x4_s2": dict(n=2, cin=64, cout=128, h=32, w=32, k=4, stride=2, pad=1),
    "wgrad_4x4_s1_cout1": dict(n=2, cin=512, cout=1, h=31, w=31, k=4, pad=1),
    "wgrad_convT3x3_s2": dict(n=2, cin=256, cout=128, h=16, w=16, k=3, stride=2, pad=1, transposed=True, out_pad=1),
    "wgrad_7x7_cout3": dict(n=1, cin=64, cout=3, h=38, w=38, k=7),
    "wgrad_rowpack8_7x7": dict(n=2, cin=3, cout=64, h=38, w=38, k=7, rowpack=8),
    "wgrad_rowpack16_4x4_s2": dict(n=2, cin=3, cout=64, h=34, w=34, k=4, stride=2, rowpack=16),
}
