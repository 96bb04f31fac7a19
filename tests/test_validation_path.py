"""Device-side validation path (new_multi/train5.py:97-110, util/util.py:51-65,
my_eval.py:52-56): prediction quantisation and the OpenCV bilinear resize must be BIT-EXACT with numpy / cv2
(the reference's own calls, restated in oracle/networks_oracle.py), the metrics within 1e-5."""
import numpy as np
import pytest
import torch

from oracle import networks_oracle as O

cv2 = pytest.importorskip("cv2")


def test_oracle_resize_is_opencv():
    """CPU: the oracle's resize IS cv2.resize; this pins the fixed-point model the kernel implements
    (documented in csrc/metrics.cu) against OpenCV on the shapes the validation loop uses."""
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (24, 40), dtype=np.uint8)
    out = cv2.resize(src, (83, 47))
    # model of OpenCV's INTER_LINEAR for uint8: 11-bit coefficients, weights clamped along x, indices along y
    def table(ssize, dsize, clamp):
        ofs, al = [], []
        dmax = dsize
        for d in range(dsize):
            f = np.float32((d + 0.5) * (ssize / dsize) - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - s)
            if clamp:
                if s < 0:
                    f, s = np.float32(0), 0
                if s + 1 >= ssize:
                    dmax = min(dmax, d)
                    if s >= ssize - 1:
                        f, s = np.float32(0), ssize - 1
            ofs.append(s)
            al.append((int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048)))),
                       int(np.rint(np.float32(f * np.float32(2048))))))
        return ofs, al, dmax
    xo, xa, xmax = table(40, 83, True)
    yo, ya, _ = table(24, 47, False)
    got = np.zeros_like(out)
    S = src.astype(np.int64)
    for dy in range(47):
        y0, y1 = min(max(yo[dy], 0), 23), min(max(yo[dy] + 1, 0), 23)
        for dx in range(83):
            sx = xo[dx]
            if dx >= xmax:
                h0, h1 = S[y0, sx] * 2048, S[y1, sx] * 2048
            else:
                x1 = min(sx + 1, 39)
                h0 = S[y0, sx] * xa[dx][0] + S[y0, x1] * xa[dx][1]
                h1 = S[y1, sx] * xa[dx][0] + S[y1, x1] * xa[dx][1]
            got[dy, dx] = min(255, max(0, (((ya[dy][0] * (h0 >> 4)) >> 16) + ((ya[dy][1] * (h1 >> 4)) >> 16) + 2) >> 2))
    assert np.array_equal(got, out)


@pytest.mark.gpu
@pytest.mark.parametrize("sh,sw,dh,dw", [(192, 640, 375, 1242), (192, 576, 375, 1242), (64, 80, 37, 123),
                                          (375, 1242, 192, 640), (31, 57, 31, 57)])
def test_resize_matches_opencv_bit_exact(sh, sw, dh, dw):
    from cycle_depth_estimation_b200 import ops
    rng = np.random.default_rng(sh * 7 + dw)
    src = rng.integers(0, 256, (3, sh, sw), dtype=np.uint8)
    got = ops.resize_linear_u8(torch.from_numpy(src).cuda(), dh, dw).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], cv2.resize(src[i], (dw, dh))), i


@pytest.mark.gpu
def test_prediction_quantisation_bit_exact():
    from cycle_depth_estimation_b200 import ops
    rng = np.random.default_rng(11)
    dep = (rng.random((4, 96, 160), dtype=np.float32) * 2 - 1) * np.array([1.0, 0.7, 0.999, 0.3], np.float32)[:, None, None]
    dep[1, :3] = -1.0
    got = ops.depth_pred_to_u8(torch.from_numpy(dep).cuda()).cpu().numpy()
    for i in range(4):
        assert np.array_equal(got[i], O.prediction_png_u8(dep[i])), i


@pytest.mark.gpu
def test_eval_metric_from_predictions_matches_the_png_round_trip():
    from cycle_depth_estimation_b200 import my_eval
    rng = np.random.default_rng(12)
    n = 5
    dep = rng.random((n, 192, 640), dtype=np.float32) * 1.9 - 0.95
    gt = rng.integers(0, 80, (n, 375, 1242), dtype=np.uint8)
    gt[rng.random((n, 375, 1242)) < 0.3] = 0
    means, per = my_eval.eval_metric_from_predictions(torch.from_numpy(dep).cuda(), gt)
    ref_means, ref_per = O.eval_metric_from_predictions(list(dep), list(gt))
    assert np.allclose(np.asarray(means, np.float64), np.asarray(ref_means, np.float64), rtol=0, atol=1e-5)
    assert np.allclose(per.astype(np.float64), ref_per[:n].astype(np.float64), rtol=0, atol=1e-5)
    # the a1/a2/a3 fractions are ratios of exact counts
    assert np.array_equal(per[:, 4:7], ref_per[:n, 4:7])
