"""K7 parity: depth metrics kernel against the numpy oracle (new_multi/my_eval.py semantics).
Bar: the three threshold COUNTS bit-exact, metric values within 1e-5 (north_star)."""
import numpy as np
import pytest
import torch

from oracle import networks_oracle as O

pytestmark = pytest.mark.gpu


def _pairs(n, h, w, seed0=2019):
    gts, preds = [], []
    for i in range(n):
        rng = np.random.default_rng(seed0 + i)
        gt = rng.integers(0, 80, (h, w), dtype=np.uint8)
        gt[rng.random((h, w)) < 0.3] = 0
        gts.append(gt)
        preds.append(rng.integers(0, 256, (h, w), dtype=np.uint8))
    return np.stack(gts), np.stack(preds)


def _check(gts, preds):
    from cycle_depth_estimation_b200 import my_eval
    got = my_eval.per_image_errors(gts, preds)
    _, ref = O.eval_metric_arrays(list(gts), list(preds))
    for i in range(len(gts)):
        gt, pred = gts[i], preds[i] / 255 * 80
        pred = np.clip(pred, 1, 50)
        mask = np.logical_and(gt > 1, gt < 50)
        r = O.compute_errors(gt[mask], pred[mask])
        cnt = int(mask.sum())
        assert int(got[i, 7]) == cnt
        for k in (4, 5, 6):  # exact counts
            assert round(got[i, k] * cnt) == round(r[k] * cnt), (i, k, got[i, k], r[k])
            assert got[i, k] == r[k], (i, k)
        for k in range(4):
            assert abs(got[i, k] - r[k]) <= 1e-5 * max(1.0, abs(r[k])), (i, k, got[i, k], r[k])
    means, _ = my_eval.eval_metric_arrays(gts, preds)
    ref_means, _ = O.eval_metric_arrays(list(gts), list(preds))
    for a, b in zip(means, ref_means):
        assert abs(float(a) - float(b)) <= 1e-5, (means, ref_means)


def test_metrics_kitti_shape():
    _check(*_pairs(6, 375, 1242))


def test_metrics_odd_small_shapes():
    _check(*_pairs(5, 37, 53, seed0=7))
    _check(*_pairs(3, 16, 16, seed0=11))


def test_metrics_narrow_prediction_range():
    gts, preds = _pairs(2, 64, 80, seed0=3)
    preds = (preds // 16 + 100).astype(np.uint8)   # few distinct values, all inside the clamp range
    _check(gts, preds)


def test_metrics_empty_mask_raises():
    from cycle_depth_estimation_b200 import my_eval
    gt = np.zeros((1, 32, 32), np.uint8)
    pred = np.full((1, 32, 32), 128, np.uint8)
    with pytest.raises(ValueError):
        my_eval.per_image_errors(gt, pred)


def test_compute_errors_single_pair_api():
    from cycle_depth_estimation_b200 import my_eval
    gts, preds = _pairs(1, 48, 64, seed0=5)
    got = my_eval.compute_errors(gts[0], preds[0])
    pred = np.clip(preds[0] / 255 * 80, 1, 50)
    mask = np.logical_and(gts[0] > 1, gts[0] < 50)
    ref = O.compute_errors(gts[0][mask], pred[mask])
    assert np.allclose(got, ref, rtol=0, atol=1e-5)
