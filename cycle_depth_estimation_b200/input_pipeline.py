"""Device side of the loaders' per-sample arithmetic: what ``datasets/dataset_synthia.py:149-208``
and ``new_multi/try_data.py:157-285`` compute with numpy / torchvision on the host between the image decode + resize
and ``set_input`` — multi-range depth labels, label-id remapping, ToTensor + Normalize — as three HBM-bound kernels on
batches that are already resident on the GPU.  Results are bit-identical to the reference statements.

``resize_bilinear_u8`` / ``resize_nearest_u8`` / ``hflip_u8`` are the loaders' ``Image.resize(size, BILINEAR)`` of the RGB
images, ``Image.resize(size, NEAREST)`` of the label images and the ``F.hflip`` of ``paired_transform`` on uint8 batches
resident on the GPU, bit-exact with Pillow (the window / coefficient / position tables are built here on the host in
double precision exactly as Pillow builds them; the kernels do the integer arithmetic).

Not covered (host side stays as in the reference): file decoding, the resize of the 16-/32-bit depth PNGs, the random
+-5 degree rotation of ``paired_transform``, the ``Canny`` edge maps.  ``cdb_resize_linear_u8`` (``ops.resize_linear_u8``)
covers OpenCV's uint8 bilinear resize.
"""
import math
import ctypes as C

import numpy as np
import torch

from . import _lib
from .ops import _p, _require_cuda, _stream

check = _lib.check


# ---------------------------------------------------------------------------------------------------------------
# label-id tables: the reference's remapping loops, composed on the host into one 256-entry table
# ---------------------------------------------------------------------------------------------------------------
def label_lut_masked(mapping):
    """datasets/dataset_synthia.py:172-183: ``copy = lab.copy(); for k, v in mapping.items(): copy[lab == k] = v`` —
    every rule tests the ORIGINAL ids, so the rules do not chain. Returns uint8[256]."""
    lab = np.arange(256, dtype=np.uint8)
    out = lab.copy()
    for k, v in mapping.items():
        out[lab == k] = v
    return out.astype(np.uint8)


def label_lut_sequential(mapping, zero_to=None):
    """new_multi/try_data.py:199-204: ``lab = lab.astype(float32); lab[lab == 0] = zero_to; for k, v in
    mapping.items(): lab[lab.copy() == k] = v`` — every rule tests the CURRENT values, so an id mapped onto a later key
    is mapped again (the reference's behaviour, kept). The final ``astype(np.uint8)`` (:224) is included."""
    lab = np.arange(256, dtype=np.float32)
    if zero_to is not None:
        lab[lab == 0] = zero_to
    for k, v in mapping.items():
        lab[lab.copy() == k] = v
    return lab.astype(np.uint8)


def label_lut_offset(offset, floor=0):
    """new_multi/try_data.py:208-211: ``lab = lab.astype(float32) + offset; lab[lab < floor] = floor`` (then uint8)."""
    lab = np.arange(256, dtype=np.float32) + np.float32(offset)
    lab[lab < floor] = floor
    return lab.astype(np.uint8)


def remap_labels(labels_u8, lut):
    """labels_u8: uint8 CUDA tensor of any shape; lut: uint8[256] (numpy or tensor). Returns the int64 class-id tensor
    ``MaskToTensor`` produces (new_multi/try_data.py:26-28), same shape."""
    _require_cuda(labels_u8)
    if labels_u8.dtype != torch.uint8:
        raise TypeError("uint8 label images expected")
    src = labels_u8.contiguous()
    lut_t = torch.as_tensor(np.asarray(lut, dtype=np.uint8) if not torch.is_tensor(lut) else lut)
    if lut_t.numel() != 256 or lut_t.dtype != torch.uint8:
        raise ValueError("a 256-entry uint8 table is expected")
    lut_d = lut_t.to(src.device)
    out = torch.empty(src.shape, dtype=torch.int64, device=src.device)
    check(_lib.lib().cdb_label_lut_i64(_p(src), C.c_int64(src.numel()), _p(lut_d), _p(out), _stream()))
    return out


# ---------------------------------------------------------------------------------------------------------------
# depth labels and image normalisation
# ---------------------------------------------------------------------------------------------------------------
def depth_labels(depth):
    """depth: fp32 CUDA [N,H,W] raw depth maps (after the loader's resize). Returns (dep_l_syn [N,1,H,W],
    depth_l_s [N,4,H,W]) as new_multi/try_data.py:240-272 builds them per sample (see cdb_depth_labels)."""
    _require_cuda(depth)
    if depth.dtype != torch.float32 or depth.dim() != 3:
        raise TypeError("fp32 [N,H,W] depth maps expected")
    d = depth.contiguous()
    n, h, w = d.shape
    L = _lib.lib()
    ws = torch.empty((max(1, L.cdb_depth_labels_workspace(n) // 4),), dtype=torch.int32, device=d.device)
    dep = torch.empty((n, 1, h, w), dtype=torch.float32, device=d.device)
    lab = torch.empty((n, 4, h, w), dtype=torch.float32, device=d.device)
    check(L.cdb_depth_labels(_p(d), n, C.c_int64(h * w), _p(dep), _p(lab), _p(ws), C.c_size_t(ws.numel() * 4), _stream()))
    return dep, lab


def normalize_images(images_u8, mean=0.5, std=0.5):
    """images_u8: uint8 CUDA [N,H,W,C] (decoded, resized RGB). Returns fp32 [N,C,H,W] =
    transforms.Normalize((mean,)*C, (std,)*C)(transforms.ToTensor()(img)) (new_multi/try_data.py:425)."""
    _require_cuda(images_u8)
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
        raise TypeError("uint8 [N,H,W,C] images expected")
    src = images_u8.contiguous()
    n, h, w, c = src.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=src.device)
    check(_lib.lib().cdb_image_normalize_u8(_p(src), n, C.c_int64(h * w), c, C.c_float(mean), C.c_float(std), _p(out),
                                            _stream()))
    return out


# ---------------------------------------------------------------------------------------------------------------
# Pillow resizes and flips (datasets/dataset_synthia.py:154-167, 228-232; new_multi/try_data.py:164-167, 377-386)
# ---------------------------------------------------------------------------------------------------------------
_PIL_PRECISION_BITS = 32 - 8 - 2


def _bilinear_filter(x):
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def pil_coeffs(in_size, out_size, support=1.0, filt=_bilinear_filter):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c) for the whole-image box: returns
    (bounds int32 [out,2] = (first source index, taps), kk int32 [out, ksize] 22-bit fixed point, ksize).  Python floats
    are C doubles and int() truncates like the C casts, so the tables equal Pillow's bit for bit."""
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    sup = support * filterscale
    ksize = int(math.ceil(sup)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - sup + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + sup + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << _PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PIL_PRECISION_BITS))
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, kk, ksize


def pil_nearest_table(in_size, out_size):
    """Source index of every output coordinate of Image.resize(size, NEAREST): ImagingScaleAffine (Geometry.c)
    pretabulates COORD(xo) with xo starting at scale * 0.5 and growing by repeated addition of scale in double."""
    scale = float(in_size) / out_size
    tab = np.full((out_size,), -1, dtype=np.int32)
    xo = 0.0 + scale * 0.5
    for x in range(out_size):
        xin = -1 if xo < 0.0 else int(xo)
        if 0 <= xin < in_size:
            tab[x] = xin
        xo += scale
    return tab


_TABLES = {}


def _dev_tables(key, build, device):
    k = (key, str(device))
    if k not in _TABLES:
        _TABLES[k] = tuple(torch.as_tensor(np.ascontiguousarray(t)).to(device) if isinstance(t, np.ndarray) else t
                           for t in build())
    return _TABLES[k]


def _u8_nhwc(images_u8):
    _require_cuda(images_u8)
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
        raise TypeError("uint8 [N,H,W,C] images expected")
    return images_u8.contiguous()


def resize_bilinear_u8(images_u8, size):
    """``Image.resize(size, Image.BILINEAR)`` (size = (width, height), as PIL takes it) of every image of a uint8 CUDA
    batch [N,H,W,C]; bit-exact with Pillow (antialiased when shrinking, as Pillow's BILINEAR is)."""
    src = _u8_nhwc(images_u8)
    n, sh, sw, c = src.shape
    dw, dh = int(size[0]), int(size[1])
    L = _lib.lib()
    bx = kx = by = ky = None
    ksx = ksy = 0
    if dw != sw:
        bx, kx, ksx = _dev_tables(("bil", sw, dw), lambda: pil_coeffs(sw, dw), src.device)
    if dh != sh:
        by, ky, ksy = _dev_tables(("bil", sh, dh), lambda: pil_coeffs(sh, dh), src.device)
    ws_bytes = L.cdb_pil_resample_workspace(n, sh, dw, c) if (dw != sw and dh != sh) else 0
    ws = torch.empty((max(1, ws_bytes),), dtype=torch.uint8, device=src.device)
    out = torch.empty((n, dh, dw, c), dtype=torch.uint8, device=src.device)
    check(L.cdb_pil_resample_u8(_p(src), n, sh, sw, c, _p(out), dh, dw, _p(bx), _p(kx), ksx, _p(by), _p(ky), ksy, _p(ws),
                                C.c_size_t(ws.numel()), _stream()))
    return out


def _gather(src, ytab, xtab):
    n, sh, sw, c = src.shape
    out = torch.empty((n, ytab.numel(), xtab.numel(), c), dtype=torch.uint8, device=src.device)
    check(_lib.lib().cdb_gather_rows_cols_u8(_p(src), n, sh, sw, c, _p(out), ytab.numel(), xtab.numel(), _p(ytab), _p(xtab),
                                             _stream()))
    return out


def resize_nearest_u8(images_u8, size):
    """``Image.resize(size, Image.NEAREST)`` of uint8 label / image batches [N,H,W,C] (size = (width, height))."""
    src = _u8_nhwc(images_u8)
    _, sh, sw, _ = src.shape
    dw, dh = int(size[0]), int(size[1])
    (xtab,) = _dev_tables(("near", sw, dw), lambda: (pil_nearest_table(sw, dw),), src.device)
    (ytab,) = _dev_tables(("near", sh, dh), lambda: (pil_nearest_table(sh, dh),), src.device)
    return _gather(src, ytab, xtab)


def hflip_u8(images_u8, flip_mask=None):
    """``F.hflip`` of paired_transform on a uint8 batch [N,H,W,C].  flip_mask: optional bool sequence of length N (the
    per-sample ``random.random() > 0.5`` decisions the caller drew, in the reference's order); None flips every image."""
    src = _u8_nhwc(images_u8)
    n, sh, sw, _ = src.shape
    (ytab,) = _dev_tables(("id", sh), lambda: (np.arange(sh, dtype=np.int32),), src.device)
    (xrev,) = _dev_tables(("rev", sw), lambda: (np.arange(sw - 1, -1, -1, dtype=np.int32),), src.device)
    if flip_mask is None:
        return _gather(src, ytab, xrev)
    mask = [bool(m) for m in flip_mask]
    if len(mask) != n:
        raise ValueError("flip_mask needs one entry per image")
    out = src.clone()
    idx = [i for i, m in enumerate(mask) if m]
    if idx:
        sel = torch.as_tensor(idx, device=src.device)
        out[sel] = _gather(src[sel].contiguous(), ytab, xrev)
    return out
