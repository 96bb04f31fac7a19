"""Drop-in for the reference's ``new_multi/my_eval.py`` (compute_errors :7-31, eval_metric :35-108) on
the K7 kernels. Inputs are the uint8 images the reference reads with ``cv2.imread(path, 0)``; arithmetic
runs on the GPU, only the 7 numbers per image come back.
"""
import os

import numpy as np
import torch

from . import ops

NAMES = ('abs_rel', 'sq_rel', 'rmse', 'rmse_log', 'a1', 'a2', 'a3')


def _as_u8_cuda(a, device):
    t = torch.as_tensor(np.ascontiguousarray(a)) if not torch.is_tensor(a) else a
    if t.dtype != torch.uint8:
        raise TypeError("uint8 images expected (cv2.imread(path, 0) semantics)")
    return t.to(device, non_blocking=True).contiguous()


def per_image_errors(gt, pred, device=None):
    """gt, pred: uint8 [n,h,w] (numpy or torch, host or device). Returns float64 numpy [n,8]:
    abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, masked pixel count — the values
    ``compute_errors(gt[mask], clip(pred/255*80,1,50)[mask])`` of the reference yields per image."""
    device = device or torch.device('cuda')
    g, p = _as_u8_cuda(gt, device), _as_u8_cuda(pred, device)
    if g.dim() == 2:
        g, p = g[None], p[None]
    out = ops.depth_metrics(g, p).cpu().numpy()
    if (out[:, 7] == 0).any():
        # numpy raises on .min() of an empty selection (new_multi/my_eval.py:10)
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    return out


def compute_errors(ground_truth, predication):
    """new_multi/my_eval.py:7-31 for one image pair given as uint8 arrays: ``ground_truth`` the depth
    image, ``predication`` the uint8 prediction; the clamp / mask of eval_metric (:56,:78-84) is applied
    inside. Returns (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3)."""
    r = per_image_errors(np.asarray(ground_truth)[None], np.asarray(predication)[None])[0]
    return tuple(float(v) for v in r[:7])


def _sharded_per_image(gts, preds, per_image_fn):
    """One process per GPU: rank r evaluates the images r, r + world, r + 2 world, ...; the float32
    per-image rows are all-gathered and put back in image order, so that every rank then performs the reference's own
    ordered float32 reduction (:108) and obtains bit-identical means whatever the number of GPUs.  No other
    collective is involved: the images are independent."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(gts)
    mine = list(range(rank, n, world))
    rows = max((n + world - 1) // world, 1)
    local = np.zeros((rows, 7), np.float32)
    if mine:
        sel = lambda a: a[mine] if (torch.is_tensor(a) or isinstance(a, np.ndarray)) else np.stack([a[i] for i in mine])
        local[:len(mine)] = np.asarray(per_image_fn(sel(gts), sel(preds)))[:, :7].astype(np.float32)
    dev = torch.device('cuda') if dist.get_backend() == 'nccl' else torch.device('cpu')
    mine_t = torch.from_numpy(local).to(dev)
    gathered = torch.empty((world * rows, 7), dtype=torch.float32, device=dev)   # concatenation form (gloo and nccl)
    dist.all_gather_into_tensor(gathered, mine_t)
    g = gathered.view(world, rows, 7).cpu().numpy()
    per = np.empty((n, 7), np.float32)
    for i in range(n):
        per[i] = g[i % world, i // world]
    return per


def eval_metric_arrays(gts, preds, per_image_fn=None):
    """eval_metric (:35-108) on in-memory stacks of equal-sized uint8 images. Per-image results are
    stored as float32 and reduced as ``float32_array.sum() / count`` like the reference (:37-43,:108).
    With ``torch.distributed`` initialised (one process per GPU) the images are sharded round-robin over the ranks
    (``_sharded_per_image``); every rank returns the same means.  ``per_image_fn`` replaces the device evaluation
    (tests of the sharding logic on CPU / gloo)."""
    import torch.distributed as dist
    fn = per_image_fn or per_image_errors
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        per = _sharded_per_image(gts, preds, fn)
    else:
        per = np.asarray(fn(gts, preds))[:, :7].astype(np.float32)
    n = per.shape[0]
    acc = np.zeros((max(n, 1000), 7), np.float32)
    acc[:n] = per
    return tuple(acc[:, k].sum() / n for k in range(7)), acc[:n]


def eval_metric(gt_p='/home/dut-ai/Documents/depth_selection/val_selection_cropped/groundtruth_depth',
                pre_p='/home/dut-ai/Documents/depth_selection/val_selection_cropped/pred'):
    """File-based entry point with the reference's semantics (reads PNG pairs with OpenCV, resizes the
    prediction to the ground-truth size). The directory defaults are the reference's hard-coded ones."""
    import cv2
    files_2 = set(os.listdir(pre_p))
    gts, preds = [], []
    for f in os.listdir(gt_p):
        if f in files_2:
            gt = cv2.imread(os.path.join(gt_p, f), 0)
            pred = cv2.imread(os.path.join(pre_p, f), 0)
            preds.append(cv2.resize(pred, (gt.shape[1], gt.shape[0])))
            gts.append(gt)
    shapes = {g.shape for g in gts}
    results = np.zeros((len(gts), 7), np.float32)
    for shp in shapes:  # one launch per image size
        idx = [i for i, g in enumerate(gts) if g.shape == shp]
        per = per_image_errors(np.stack([gts[i] for i in idx]), np.stack([preds[i] for i in idx]))[:, :7]
        results[idx] = per.astype(np.float32)
    n = len(gts)
    acc = np.zeros((max(n, 1000), 7), np.float32)
    acc[:n] = results
    return tuple(acc[:, k].sum() / n for k in range(7))


def eval_metric_from_predictions(pred_dep, gts):
    """The validation loop of new_multi/train5.py:85-110 without its PNG round trip: ``pred_dep`` is the network's
    depth output (fp32 CUDA [n,h,w], [-1,1] convention, ``real_dep_ref``), ``gts`` the uint8 ground-truth images
    [n,H,W].  Quantisation (tensor2im + max-normalisation + PNG rounding), cv2.resize to the ground-truth size
    (bit-exact INTER_LINEAR) and the metrics all run on the device.  Returns (7 float32 means, per-image [n,7])."""
    device = pred_dep.device
    g = _as_u8_cuda(gts, device)
    if g.dim() == 2:
        g = g[None]
    p8 = ops.depth_pred_to_u8(pred_dep.detach().contiguous().float())
    if tuple(p8.shape[1:]) != tuple(g.shape[1:]):
        p8 = ops.resize_linear_u8(p8, int(g.shape[1]), int(g.shape[2]))
    return eval_metric_arrays(g, p8)
