"""Generate golden vectors for the episode-loader transform pipeline from the REAL libraries the reference calls
(PIL + torchvision, data/datamgr.py:37-46, data/additional_transforms.py:19-34), on seeded synthetic images and
explicit augmentation parameters.  The fixtures are committed (tests/golden/transforms.npz); tests/test_transforms_oracle.py
requires oracle/transforms.py to reproduce them bit for bit.

``RandomSizedCrop`` / ``Scale`` are the pre-0.2 names of ``RandomResizedCrop`` / ``Resize``; the parameter-explicit
functional forms are used here (F.resized_crop, F.resize, F.center_crop, F.hflip, F.to_tensor, F.normalize) together
with the reference's own ``ImageJitter`` loop body driven by given factors.

Run:  python tests/golden/make_golden_transforms.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

MEAN = [0.485, 0.456, 0.406]
STD = [0.229, 0.224, 0.225]

# (seed, H, W, out size) -- shapes like CUB / miniImagenet files, plus small / up-scaling / extreme cases
AUG_CASES = [(1, 375, 500, 84), (2, 333, 500, 84), (3, 84, 84, 84), (4, 40, 57, 84), (5, 120, 97, 32),
             (6, 500, 375, 224), (7, 200, 300, 84), (8, 64, 64, 84)]
PLAIN_CASES = [(11, 375, 500, 84), (12, 90, 130, 84), (13, 96, 96, 84), (14, 500, 333, 224), (15, 70, 50, 32)]


def synth_image(seed, h, w):
    """Smooth structure + noise, full 0..255 range (so the clipping branches of blend are exercised)."""
    r = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.empty((h, w, 3), np.float64)
    for c in range(3):
        img[..., c] = 127 + 90 * np.sin(xx / (7.0 + 3 * c) + seed) * np.cos(yy / (11.0 - 2 * c)) + r.randn(h, w) * 40
    return np.clip(img, 0, 255).astype(np.uint8)


def crop_for(seed, h, w):
    r = np.random.RandomState(1000 + seed)
    ch = int(r.randint(max(1, h // 4), h + 1))
    cw = int(r.randint(max(1, w // 4), w + 1))
    top = int(r.randint(0, h - ch + 1))
    left = int(r.randint(0, w - cw + 1))
    return top, left, ch, cw


def factors_for(seed):
    g = torch.Generator().manual_seed(2000 + seed)
    rt = torch.rand(3, generator=g)
    return [float(0.4 * (rt[i] * 2.0 - 1.0) + 1) for i in range(3)]       # additional_transforms.py:30, fp32 tensor math


def main():
    from PIL import Image, ImageEnhance
    import torchvision.transforms.functional as F
    from torchvision.transforms import InterpolationMode

    out = {}
    for seed, h, w, size in AUG_CASES:
        img = synth_image(seed, h, w)
        top, left, ch, cw = crop_for(seed, h, w)
        if seed == 3:
            top, left, ch, cw = 0, 0, h, w          # identity-size crop: both passes skipped in Pillow
        fac = factors_for(seed)
        if seed == 8:
            fac = [1.0, 0.0, 1.4]                    # the alpha == 1 / alpha == 0 short-cuts and a clipping factor
        flip = seed % 2
        pil = Image.fromarray(img)
        o = F.resized_crop(pil, top, left, ch, cw, [size, size], InterpolationMode.BILINEAR)
        out["aug%d_resized" % seed] = np.array(o)
        for enh, f in zip((ImageEnhance.Brightness, ImageEnhance.Contrast, ImageEnhance.Color), fac):
            o = enh(o).enhance(torch.tensor(f, dtype=torch.float32)).convert("RGB")     # the reference passes a tensor
        out["aug%d_jitter" % seed] = np.array(o)
        if flip:
            o = F.hflip(o)
        t = F.normalize(F.to_tensor(o), MEAN, STD)
        out["aug%d_out" % seed] = t.numpy()
        out["aug%d_params" % seed] = np.array([top, left, ch, cw, flip], np.int64)
        out["aug%d_factors" % seed] = np.array(fac, np.float32)
    for seed, h, w, size in PLAIN_CASES:
        img = synth_image(seed, h, w)
        pil = Image.fromarray(img)
        big = int(size * 1.15)
        o = F.center_crop(F.resize(pil, [big, big], InterpolationMode.BILINEAR), [size, size])
        out["plain%d_out" % seed] = F.normalize(F.to_tensor(o), MEAN, STD).numpy()
    np.savez_compressed(os.path.join(HERE, "transforms.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
