"""the loaders' per-sample arithmetic.

CPU part (always): the numpy restatement (oracle/input_pipeline_oracle.py) and the host-side label tables
(input_pipeline.label_lut_*) against tests/golden/input_pipeline.pt, which oracle/make_golden.py produced by EXECUTING
the reference's own source lines (new_multi/try_data.py:199-211,240-272; datasets/dataset_synthia.py:172-175;
torchvision ToTensor + Normalize) on seeded arrays — bit for bit, NaN patterns of degenerate depth ranges included.
GPU part: the three kernels against the oracle, bit-exact, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import input_pipeline_oracle as OI

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLD, "input_pipeline.pt"), weights_only=False)


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def test_oracle_depth_labels_match_the_reference_lines(fx):
    kinds = set()
    for case in fx['depth']:
        dep, lab = OI.depth_labels(case['depth'])
        assert _same(dep, case['dep_l']) and _same(lab, case['depth_l_s'])
        kinds.add((bool(np.isnan(case['depth_l_s']).any()), bool(np.isnan(case['dep_l']).any())))
    assert (True, False) in kinds and (False, False) in kinds and (True, True) in kinds   # degenerate ranges are covered
    # the reference's quirk (:266): the last label subtracts the minimum of the NORMALISED fourth range (-1)
    c = fx['depth'][0]
    d5 = np.minimum(c['depth'], np.float32(2000))
    want = 2 * (d5 + np.float32(1)) / (d5.max() - d5.min()) - 1
    assert _same(want.astype(np.float32), c['depth_l_s'][3])


def test_label_tables_match_the_reference_lines(fx):
    from cycle_depth_estimation_b200 import input_pipeline as IP
    s = fx['labels']['sequential']
    lut = IP.label_lut_sequential(s['mapping'], zero_to=s['zero_to'])
    assert lut.dtype == np.uint8 and lut.shape == (256,)
    assert np.array_equal(lut[s['lab']], s['out'])
    assert torch.equal(OI.remap_sequential(s['lab'], s['mapping'], s['zero_to']), torch.from_numpy(s['out']).long())
    assert np.array_equal(IP.label_lut_offset(-6, 0)[s['lab']], s['out_target'])
    assert torch.equal(OI.remap_offset(s['lab']), torch.from_numpy(s['out_target']).long())
    m = fx['labels']['masked']
    assert np.array_equal(IP.label_lut_masked(m['mapping'])[m['lab']], m['out'])
    assert torch.equal(OI.remap_masked(m['lab'], m['mapping']), torch.from_numpy(m['out']).long())
    # sequential rules chain (an id mapped onto a later key is mapped again), masked rules do not
    chain = {1: 2, 2: 3}
    assert IP.label_lut_sequential(chain)[1] == 3 and IP.label_lut_masked(chain)[1] == 2


def test_oracle_normalize_matches_torchvision(fx):
    n = fx['normalize']
    assert torch.equal(OI.normalize(n['img']), n['out'])


def test_host_side_refuses_cpu_tensors():
    from cycle_depth_estimation_b200 import input_pipeline as IP
    with pytest.raises(RuntimeError):
        IP.depth_labels(torch.zeros(1, 4, 4))
    with pytest.raises(RuntimeError):
        IP.remap_labels(torch.zeros(4, 4, dtype=torch.uint8), np.arange(256, dtype=np.uint8))
    with pytest.raises(RuntimeError):
        IP.normalize_images(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_depth_labels_kernel_bit_exact(fx):
    from cycle_depth_estimation_b200 import input_pipeline as IP
    for case in fx['depth']:
        dep, lab = IP.depth_labels(torch.from_numpy(case['depth']).cuda()[None])
        assert _same(dep[0].cpu().numpy(), case['dep_l']) and _same(lab[0].cpu().numpy(), case['depth_l_s'])
    # a batch at the loader's size (192 x 640) with per-image ranges, against the oracle
    rng = np.random.default_rng(5)
    batch = np.stack([rng.uniform(0, hi, (192, 640)).astype(np.float32) for hi in (65535, 9000, 3000, 700)])
    batch[1, :10] = np.float32(8000.5)
    dep, lab = IP.depth_labels(torch.from_numpy(batch).cuda())
    for i in range(batch.shape[0]):
        d0, l0 = OI.depth_labels(batch[i])
        assert _same(dep[i].cpu().numpy(), d0), i
        assert _same(lab[i].cpu().numpy(), l0), i


@pytest.mark.gpu
def test_label_remap_kernel_bit_exact(fx):
    from cycle_depth_estimation_b200 import input_pipeline as IP
    s = fx['labels']['sequential']
    rng = np.random.default_rng(6)
    lab = rng.integers(0, 256, (3, 192, 640), dtype=np.uint8)
    out = IP.remap_labels(torch.from_numpy(lab).cuda(), IP.label_lut_sequential(s['mapping'], zero_to=7))
    assert out.dtype == torch.int64 and tuple(out.shape) == (3, 192, 640)
    for i in range(3):
        assert torch.equal(out[i].cpu(), OI.remap_sequential(lab[i], s['mapping'], 7))
    out = IP.remap_labels(torch.from_numpy(lab).cuda(), IP.label_lut_offset(-6, 0))
    assert torch.equal(out[1].cpu(), OI.remap_offset(lab[1]))
    m = fx['labels']['masked']
    out = IP.remap_labels(torch.from_numpy(lab).cuda(), IP.label_lut_masked(m['mapping']))
    assert torch.equal(out[2].cpu(), OI.remap_masked(lab[2], m['mapping']))


@pytest.mark.gpu
def test_image_normalize_kernel_bit_exact(fx):
    from cycle_depth_estimation_b200 import input_pipeline as IP
    n = fx['normalize']
    out = IP.normalize_images(torch.from_numpy(n['img']).cuda()[None])
    assert torch.equal(out[0].cpu(), n['out'])
    rng = np.random.default_rng(7)
    imgs = rng.integers(0, 256, (4, 192, 640, 3), dtype=np.uint8)
    imgs[0, 0, :256, 0] = np.arange(256, dtype=np.uint8)      # every byte value
    out = IP.normalize_images(torch.from_numpy(imgs).cuda())
    for i in range(4):
        assert torch.equal(out[i].cpu(), OI.normalize(imgs[i])), i
    assert float(out.min()) >= -1.0 and float(out.max()) <= 1.0


# ---------------------------------------------------------------------------------------------------------------
# Pillow resizes / flip (datasets/dataset_synthia.py:154-167, 228-232; new_multi/try_data.py:164-167)
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def pil_fx():
    return torch.load(os.path.join(GOLD, "pil_resize.pt"), weights_only=False)


def _case_inputs(case):
    if 'img' in case:
        return case['img'], case['lab'], (1, 1)
    h, w = case['shape']
    r = np.random.default_rng(case['seed'])
    return r.integers(0, 256, (h, w, 3), dtype=np.uint8), r.integers(0, 34, (h, w), dtype=np.uint8), case['sub']


def test_oracle_pil_resize_matches_pillow_fixtures(pil_fx):
    """The numpy restatement against outputs of Pillow itself (oracle/make_golden.py::make_pil_resize), bit for bit:
    antialiased down-scaling at the datasets' aspect ratios, up-scaling, one-axis resizes, the full 1280x760 -> 640x192."""
    assert len(pil_fx['cases']) >= 7
    for case in pil_fx['cases']:
        img, lab, (sy, sx) = _case_inputs(case)
        bil = OI.pil_resize_bilinear(img, case['size'])
        assert _same(bil[::sy, ::sx], case['bilinear']), case['shape']
        assert _same(OI.pil_resize_nearest(lab, case['size'])[::sy, ::sx], case['nearest']), case['shape']
        assert _same(OI.hflip(bil)[::sy, ::sx], case['bilinear_flipped']), case['shape']


def test_oracle_pil_resize_matches_the_installed_pillow():
    """Live pin where Pillow is importable (it is in this image): KITTI-sized input at both reference target sizes."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (375, 1242, 3), dtype=np.uint8)
    lab = rng.integers(0, 34, (375, 1242), dtype=np.uint8)
    for size in ([640, 192], [576, 192]):
        assert _same(OI.pil_resize_bilinear(img, size), np.array(Image.fromarray(img).resize(size, Image.BILINEAR)))
        assert _same(OI.pil_resize_nearest(lab, size), np.array(Image.fromarray(lab).resize(size, Image.NEAREST)))


def test_host_tables_equal_the_oracle_windows():
    """input_pipeline.pil_coeffs / pil_nearest_table (what the kernels consume) against the oracle's own construction."""
    from cycle_depth_estimation_b200 import input_pipeline as IP
    for (i, o) in [(1280, 640), (760, 192), (1242, 576), (375, 192), (47, 90), (64, 64), (5, 3), (3, 7)]:
        bounds, kk, ksize = IP.pil_coeffs(i, o)
        wins = OI._pil_windows(i, o)
        assert bounds.shape == (o, 2) and kk.shape == (o, ksize)
        for xx, (xmin, k) in enumerate(wins):
            assert bounds[xx, 0] == xmin and bounds[xx, 1] == len(k)
            assert np.array_equal(kk[xx, :len(k)], k) and not kk[xx, len(k):].any()
        assert np.array_equal(IP.pil_nearest_table(i, o), OI._pil_nearest_index(i, o))


@pytest.mark.gpu
def test_pil_resize_kernels_bit_exact(pil_fx):
    from cycle_depth_estimation_b200 import input_pipeline as IP
    for case in pil_fx['cases']:
        img, lab, (sy, sx) = _case_inputs(case)
        d_img = torch.from_numpy(img).cuda()[None]
        d_lab = torch.from_numpy(lab).cuda()[None, :, :, None]
        bil = IP.resize_bilinear_u8(d_img, case['size'])
        assert _same(bil[0].cpu().numpy()[::sy, ::sx], case['bilinear']), case['shape']
        near = IP.resize_nearest_u8(d_lab, case['size'])
        assert _same(near[0, :, :, 0].cpu().numpy()[::sy, ::sx], case['nearest']), case['shape']
        assert _same(IP.hflip_u8(bil)[0].cpu().numpy()[::sy, ::sx], case['bilinear_flipped']), case['shape']


@pytest.mark.gpu
def test_pil_resize_kernels_on_batches_at_the_dataset_sizes():
    """KITTI 1242x375 -> 640x192 and 576x192, batch of 3 against the oracle per image; per-sample flip decisions."""
    from cycle_depth_estimation_b200 import input_pipeline as IP
    rng = np.random.default_rng(11)
    imgs = rng.integers(0, 256, (3, 375, 1242, 3), dtype=np.uint8)
    labs = rng.integers(0, 34, (3, 375, 1242, 1), dtype=np.uint8)
    d_imgs, d_labs = torch.from_numpy(imgs).cuda(), torch.from_numpy(labs).cuda()
    for size in ((640, 192), (576, 192)):
        bil = IP.resize_bilinear_u8(d_imgs, size).cpu().numpy()
        near = IP.resize_nearest_u8(d_labs, size).cpu().numpy()
        for i in range(3):
            assert _same(bil[i], OI.pil_resize_bilinear(imgs[i], size))
            assert _same(near[i, :, :, 0], OI.pil_resize_nearest(labs[i, :, :, 0], size))
    flipped = IP.hflip_u8(d_imgs, [True, False, True]).cpu().numpy()
    assert _same(flipped[0], OI.hflip(imgs[0])) and _same(flipped[1], imgs[1]) and _same(flipped[2], OI.hflip(imgs[2]))
    one_axis = IP.resize_bilinear_u8(d_imgs, (1242, 192)).cpu().numpy()       # vertical pass only
    assert _same(one_axis[1], OI.pil_resize_bilinear(imgs[1], (1242, 192)))
