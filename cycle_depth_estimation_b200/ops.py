"""Tensor-level wrappers over the C ABI (include/cdb200.h).

Every function takes CUDA tensors, passes raw device pointers plus geometry to libcdb200.so and
launches on torch's current stream. Nothing here computes with torch ops.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, BF16, F32, CdbAct, CdbConvGeom,
                   CdbEpilogue, CdbOut, check)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cdb200 ops need CUDA tensors (there is no CPU path)")


def round_up(a, b):
    return (a + b - 1) // b * b


def _dtype_code(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError("unsupported dtype %s" % t.dtype)


def act_view(t):
    """CdbAct for an NHWC tensor view [N,H,W,C] with unit channel stride."""
    assert t.dim() == 4 and t.stride(3) == 1, "NHWC view with contiguous channels expected"
    return CdbAct(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3], t.stride(0), t.stride(1),
                  t.stride(2), _dtype_code(t), 0)


def out_view_nhwc(t, c_real):
    """CdbOut writing c_real channels (zero for the rest) into an NHWC tensor view [N,P,Q,Cstore]."""
    assert t.dim() == 4 and t.stride(3) == 1
    return CdbOut(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], c_real, t.shape[3], _dtype_code(t),
                  t.stride(0), t.stride(1), t.stride(2), 1)


def out_view_nchw(t):
    """CdbOut writing into an NCHW tensor [N,C,P,Q] (any strides)."""
    assert t.dim() == 4
    return CdbOut(t.data_ptr(), t.shape[0], t.shape[2], t.shape[3], t.shape[1], t.shape[1], _dtype_code(t),
                  t.stride(0), t.stride(2), t.stride(3), t.stride(1))


def geom(r, s, stride=1, pad_h=0, pad_w=0, dil=1, transposed=False, rowpack=0):
    return CdbConvGeom(r, s, stride, pad_h, pad_w, dil, 1 if transposed else 0, rowpack)


def packed_weight_shape(d0, d1, r, s, rows_are_dim0, rowpack=0):
    rows = d0 if rows_are_dim0 else d1
    kdim = d1 if rows_are_dim0 else d0
    rows_pad = round_up(rows, 16)
    if rowpack:
        return rows_pad, 64, r
    return rows_pad, round_up(kdim, 64), r * s


def pack_conv_weight(w4, rows_are_dim0, rowpack=0, out=None):
    """fp32 [d0,d1,R,S] -> bf16 packed GEMM operand [rows_pad, taps*kpad]; returns (packed, rows_pad, kpad)."""
    _require_cuda(w4)
    w4 = w4.detach()
    assert w4.dtype == torch.float32 and w4.is_contiguous()
    d0, d1, r, s = w4.shape
    rows_pad, kpad, taps = packed_weight_shape(d0, d1, r, s, rows_are_dim0, rowpack)
    if out is None:
        out = torch.empty((rows_pad, taps * kpad), dtype=torch.bfloat16, device=w4.device)
    check(_lib.lib().cdb_pack_conv_weight(C.c_void_p(w4.data_ptr()), d0, d1, r, s, 1 if rows_are_dim0 else 0,
                                          rowpack, C.c_void_p(out.data_ptr()), _stream()))
    return out, rows_pad, kpad


def conv2d_fwd(g, x, wpacked, rows_pad, kpad, out, bias=None, act=ACT_NONE, slope=0.0, stats=None):
    """x: NHWC bf16 view; out: CdbOut. See cdb_conv2d_fwd."""
    _require_cuda(x, wpacked)
    xv = act_view(x)
    ep = CdbEpilogue(bias.data_ptr() if bias is not None else None, act, slope,
                     stats.data_ptr() if stats is not None else None, 0)
    check(_lib.lib().cdb_conv2d_fwd(C.byref(g), C.byref(xv), C.c_void_p(wpacked.data_ptr()), rows_pad, kpad,
                                    C.byref(out), C.byref(ep), _stream()))


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def conv2d_wgrad(g, x, dy, dw4, accumulate=False):
    """dw4 (fp32 [d0,d1,R,S]) (+)= filter gradient. x, dy: NHWC bf16 views."""
    _require_cuda(x, dy, dw4)
    assert dw4.dtype == torch.float32 and dw4.is_contiguous()
    xv, dyv = act_view(x), act_view(dy)
    L = _lib.lib()
    need = L.cdb_conv2d_wgrad_workspace(C.byref(g), C.byref(xv), C.byref(dyv))
    ws = _workspace(need, x.device)
    check(L.cdb_conv2d_wgrad(C.byref(g), C.byref(xv), C.byref(dyv), C.c_void_p(dw4.data_ptr()), dw4.shape[0],
                             dw4.shape[1], 1 if accumulate else 0, C.c_void_p(ws.data_ptr()),
                             C.c_size_t(ws.numel()), _stream()))
