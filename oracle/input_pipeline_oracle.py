"""ORACLE — test infrastructure only (imported by tests/ and oracle/make_golden.py; never by the product package).

numpy / torch-CPU restatement of the per-sample arithmetic of the reference's loaders:

* ``depth_labels``      new_multi/try_data.py:235-272 (float32 numpy, per image)
* ``remap_sequential``  new_multi/try_data.py:199-204 (+ the uint8 cast of :224 and MaskToTensor :26-28)
* ``remap_masked``      datasets/dataset_synthia.py:172-183
* ``remap_offset``      new_multi/try_data.py:208-211
* ``normalize``         transforms.ToTensor() + transforms.Normalize((0.5,)*3, (0.5,)*3) (new_multi/try_data.py:425)
* ``pil_resize_bilinear`` / ``pil_resize_nearest`` / ``hflip``   ``Image.resize([640, 192], Image.BILINEAR)`` /
  ``Image.resize(..., Image.NEAREST)`` (datasets/dataset_synthia.py:154-167, new_multi/try_data.py:164-167) and the
  ``F.hflip`` of paired_transform (:228-232).  The arithmetic lives in a third-party dependency that is not vendored in
  the reference (Pillow; ``requirements.txt`` does not pin it — the image has Pillow 12.2.0): restated here from its
  published algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal /
  Vertical_8bpc; Geometry.c: ImagingScaleAffine) and PINNED by fixtures that oracle/make_golden.py generates by calling
  Pillow itself at the reference's call-site arguments (tests/golden/pil_resize.pt), plus a live comparison with the
  installed Pillow in tests/test_input_pipeline.py.

Parity pin: the reference has no tests or golden vectors; these statements live inline in ``__getitem__`` and need the
datasets on disk.  oracle/make_golden.py therefore EXECUTES the reference's own source lines (read from
/root/reference at generation time, never copied into this repository) on seeded arrays and stores inputs + outputs
in tests/golden/input_pipeline.pt; tests/test_input_pipeline.py checks this restatement against them bit for bit.
"""
import numpy as np
import torch


def _pm1(x):
    return 2 * (x - x.min()) / (x.max() - x.min()) - 1


def depth_labels(depth):
    """depth: float32 [H,W]. Returns (dep_l [1,H,W], depth_l_s [4,H,W]) float32."""
    with np.errstate(all='ignore'):
        d = np.array(depth, dtype=np.float32)
        d2, d3, d4, d5 = d.copy(), d.copy(), d.copy(), d.copy()
        d[d > 8000] = 8000
        d2[d2 > 8000] = 8000
        d2[d < 5000] = 5000
        d2 = _pm1(d2)[None]
        d3[d > 6000] = 6000
        d3[d < 3000] = 3000
        d3 = _pm1(d3)[None]
        d4[d > 4000] = 4000
        d4[d < 1000] = 1000
        d4 = _pm1(d4)[None]
        d5[d > 2000] = 2000
        d5 = (2 * (d5 - d4.min()) / (d5.max() - d5.min()) - 1)[None]     # sic: the minimum of the NORMALISED d4
        dep = _pm1(d)[None]
        return dep, np.concatenate([d2, d3, d4, d5], axis=0)


def remap_sequential(lab_u8, mapping, zero_to=7):
    lab = np.array(lab_u8).astype(np.float32)
    lab[lab == 0] = zero_to
    for k, v in mapping.items():
        lab[lab.copy() == k] = v
    return torch.from_numpy(np.array(lab.astype(np.uint8), dtype=np.int32)).long()


def remap_masked(lab_u8, mapping):
    lab = np.array(lab_u8)
    out = lab.copy()
    for k, v in mapping.items():
        out[lab == k] = v
    return torch.from_numpy(np.array(out.astype(np.uint8), dtype=np.int32)).long()


def remap_offset(lab_u8, offset=-6, floor=0):
    lab = np.array(lab_u8).astype(np.float32)
    lab = lab + offset
    lab[lab < floor] = floor
    return torch.from_numpy(np.array(lab.astype(np.uint8), dtype=np.int32)).long()


def normalize(img_u8_hwc, mean=0.5, std=0.5):
    """uint8 [H,W,C] -> float32 [C,H,W] with torch CPU arithmetic (what torchvision's ToTensor / Normalize execute)."""
    t = torch.from_numpy(np.ascontiguousarray(img_u8_hwc)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    c = t.shape[0]
    m = torch.as_tensor((mean,) * c, dtype=torch.float32).view(-1, 1, 1)
    s = torch.as_tensor((std,) * c, dtype=torch.float32).view(-1, 1, 1)
    return t.sub_(m).div_(s)


# ---------------------------------------------------------------------------------------------------------------
# Pillow: Image.resize(size, BILINEAR / NEAREST), FLIP_LEFT_RIGHT
# ---------------------------------------------------------------------------------------------------------------
_PRECISION_BITS = 32 - 8 - 2


def _pil_windows(in_size, out_size):
    """precompute_coeffs for the bilinear (triangle, support 1) filter over the whole-image box, then
    normalize_coeffs_8bpc: list of (xmin, int coefficients) per output coordinate.  Scalar double arithmetic in the order
    of the C code (Python floats are IEEE doubles; int() truncates like a C cast)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ss = 1.0 / filterscale
    out = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ws, ww = [], 0.0
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - t if t < 1.0 else 0.0
            ws.append(w)
            ww += w
        ks = []
        for w in ws:
            if ww != 0.0:
                w = w / ww
            ks.append(int(-0.5 + w * (1 << _PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << _PRECISION_BITS)))
        out.append((xmin, np.asarray(ks, dtype=np.int64)))
    return out


def _clip8(ss):
    return np.clip(ss >> _PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bilinear(img_u8, size):
    """img_u8: uint8 [H,W,C] (or [H,W]); size = (width, height). Horizontal pass, byte rounding, vertical pass — each
    pass only when that size changes (ImagingResample)."""
    img = np.asarray(img_u8)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    dw, dh = int(size[0]), int(size[1])
    h, w, c = img.shape
    if dw != w:
        o = np.empty((h, dw, c), dtype=np.uint8)
        for xx, (xmin, k) in enumerate(_pil_windows(w, dw)):
            acc = (img[:, xmin:xmin + len(k), :].astype(np.int64) * k[None, :, None]).sum(axis=1)
            o[:, xx, :] = _clip8(acc + (1 << (_PRECISION_BITS - 1)))
        img = o
    if dh != h:
        o = np.empty((dh, img.shape[1], c), dtype=np.uint8)
        for yy, (ymin, k) in enumerate(_pil_windows(h, dh)):
            acc = (img[ymin:ymin + len(k)].astype(np.int64) * k[:, None, None]).sum(axis=0)
            o[yy] = _clip8(acc + (1 << (_PRECISION_BITS - 1)))
        img = o
    return img[:, :, 0] if squeeze else img


def _pil_nearest_index(in_size, out_size):
    scale = float(in_size) / out_size
    xo = scale * 0.5
    idx = []
    for _ in range(out_size):
        idx.append(int(xo))          # COORD(): xo is never negative for a whole-image box
        xo += scale                  # accumulated, as ImagingScaleAffine does (not x * scale)
    return np.asarray(idx, dtype=np.int64)


def pil_resize_nearest(img_u8, size):
    img = np.asarray(img_u8)
    dw, dh = int(size[0]), int(size[1])
    return img[_pil_nearest_index(img.shape[0], dh)][:, _pil_nearest_index(img.shape[1], dw)]


def hflip(img_u8):
    return np.ascontiguousarray(np.asarray(img_u8)[:, ::-1])
