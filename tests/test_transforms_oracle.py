"""The transform oracle (oracle/transforms.py) against the golden vectors produced by the real PIL / torchvision
(tests/golden/make_golden_transforms.py), bit for bit, and -- when those libraries are importable -- against them live
on fresh random cases.  No GPU."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_transforms as mg  # noqa: E402
from oracle import transforms as ot  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "transforms.npz"))


@pytest.mark.parametrize("case", mg.AUG_CASES)
def test_aug_golden(case):
    seed, h, w, size = case
    img = mg.synth_image(seed, h, w)
    top, left, ch, cw, flip = (int(v) for v in GOLD["aug%d_params" % seed])
    fac = GOLD["aug%d_factors" % seed]
    r = ot.resized_crop(img, top, left, ch, cw, size)
    assert np.array_equal(r, GOLD["aug%d_resized" % seed])
    assert np.array_equal(ot.image_jitter(r, fac), GOLD["aug%d_jitter" % seed])
    out = ot.transform_aug(img, (top, left, ch, cw), fac, flip, size)
    assert out.dtype == np.float32 and np.array_equal(out, GOLD["aug%d_out" % seed])


@pytest.mark.parametrize("case", mg.PLAIN_CASES)
def test_plain_golden(case):
    seed, h, w, size = case
    out = ot.transform_plain(mg.synth_image(seed, h, w), size)
    assert np.array_equal(out, GOLD["plain%d_out" % seed])


def test_live_libraries():
    """Fresh cases against the installed PIL / torchvision (skipped where they are absent)."""
    Image = pytest.importorskip("PIL.Image")
    ImageEnhance = pytest.importorskip("PIL.ImageEnhance")
    F = pytest.importorskip("torchvision.transforms.functional")
    from torchvision.transforms import InterpolationMode
    rs = np.random.RandomState(5)
    for t in range(12):
        h, w = int(rs.randint(8, 200)), int(rs.randint(8, 200))
        size = int(rs.choice([7, 16, 33, 84]))
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ch, cw = int(rs.randint(1, h + 1)), int(rs.randint(1, w + 1))
        top, left = int(rs.randint(0, h - ch + 1)), int(rs.randint(0, w - cw + 1))
        fac = (0.6 + 0.8 * rs.rand(3)).astype(np.float32)
        o = F.resized_crop(Image.fromarray(img), top, left, ch, cw, [size, size], InterpolationMode.BILINEAR)
        for enh, f in zip((ImageEnhance.Brightness, ImageEnhance.Contrast, ImageEnhance.Color), fac):
            o = enh(o).enhance(torch.tensor(f)).convert("RGB")
        if t % 2:
            o = F.hflip(o)
        ref = F.normalize(F.to_tensor(o), mg.MEAN, mg.STD).numpy()
        assert np.array_equal(ot.transform_aug(img, (top, left, ch, cw), fac, t % 2, size), ref)
        if min(h, w) >= 8:
            big = int(size * 1.15)
            p = F.center_crop(F.resize(Image.fromarray(img), [big, big], InterpolationMode.BILINEAR), [size, size])
            assert np.array_equal(ot.transform_plain(img, size), F.normalize(F.to_tensor(p), mg.MEAN, mg.STD).numpy())


def test_crop_sampler_matches_torchvision_logic():
    """The vectorised crop sampler of the product obeys the constraints of torchvision's get_params (restated in
    oracle.random_resized_crop_params) and draws from the same distribution."""
    from deep_kernel_transfer_b200.episode_feed import draw_crops
    g = torch.Generator().manual_seed(0)
    n = 4000
    H = np.full(n, 375)
    W = np.full(n, 500)
    top, left, ch, cw = draw_crops(H, W, g)
    assert (ch > 0).all() and (cw > 0).all() and (top >= 0).all() and (left >= 0).all()
    assert (top + ch <= H).all() and (left + cw <= W).all()
    g2 = torch.Generator().manual_seed(1)
    ref = np.array([ot.random_resized_crop_params(375, 500, g2) for _ in range(n)])
    frac, frac_ref = (ch * cw) / (375.0 * 500.0), (ref[:, 2] * ref[:, 3]) / (375.0 * 500.0)
    assert frac.min() > 0.07 and frac.max() <= 1.0
    assert abs(frac.mean() - frac_ref.mean()) < 0.02 and abs(np.log(cw / ch).mean() - np.log(ref[:, 3] / ref[:, 2]).mean()) < 0.03
    # a strip no proposal fits: the ratio-clamped centre crop
    t2, l2, h2, w2 = draw_crops(np.array([10]), np.array([1000]), g)
    assert (int(h2[0]), int(w2[0])) == (10, 13) and int(l2[0]) == (1000 - 13) // 2 and int(t2[0]) == 0
    assert ot.random_resized_crop_params(10, 1000, g2)[2:] == (10, 13)
