"""Device side of the loaders' per-sample arithmetic: what ``datasets/dataset_synthia.py:149-208``
and ``new_multi/try_data.py:157-285`` compute with numpy / torchvision on the host between the image decode + resize
and ``set_input`` — multi-range depth labels, label-id remapping, ToTensor + Normalize — as three HBM-bound kernels on
batches that are already resident on the GPU.  Results are bit-identical to the reference statements.

Not covered (host side stays as in the reference): file decoding, the PIL / OpenCV resizes of float images and the
random crop / flip augmentation (``paired_transform``), the ``Canny`` edge maps.  ``cdb_resize_linear_u8``
(``ops.resize_linear_u8``) covers OpenCV's uint8 bilinear resize.
"""
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
