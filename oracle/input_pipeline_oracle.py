"""ORACLE — test infrastructure only (imported by tests/ and oracle/make_golden.py; never by the product package).

numpy / torch-CPU restatement of the per-sample arithmetic of the reference's loaders:

* ``depth_labels``      new_multi/try_data.py:235-272 (float32 numpy, per image)
* ``remap_sequential``  new_multi/try_data.py:199-204 (+ the uint8 cast of :224 and MaskToTensor :26-28)
* ``remap_masked``      datasets/dataset_synthia.py:172-183
* ``remap_offset``      new_multi/try_data.py:208-211
* ``normalize``         transforms.ToTensor() + transforms.Normalize((0.5,)*3, (0.5,)*3) (new_multi/try_data.py:425)

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
