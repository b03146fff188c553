"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The per-image transform pipeline of the reference's episode
loader, restated in numpy integer / float32 arithmetic.

The reference composes (data/datamgr.py:37-46)
    aug:      RandomSizedCrop(S) -> ImageJitter(Brightness .4, Contrast .4, Color .4) -> RandomHorizontalFlip
              -> ToTensor -> Normalize(mean, std)
    no aug:   Scale([int(1.15 S)] * 2) -> CenterCrop(S) -> ToTensor -> Normalize(mean, std)
out of torchvision transforms on PIL images (``RandomSizedCrop`` / ``Scale`` are the old names of
``RandomResizedCrop`` / ``Resize``) and ``data/additional_transforms.py:19-34`` (PIL ``ImageEnhance``).  The arithmetic
therefore lives in two un-vendored third-party packages; what is restated here is their published algorithm:

  * Pillow ``ImagingResample`` for 8-bit images (src/libImaging/Resample.c: ``precompute_coeffs``,
    ``normalize_coeffs_8bpc``, ``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``): bilinear filter with the support
    stretched by the down-scaling factor, double-precision coefficients normalised per output pixel, converted to
    22-bit fixed point, horizontal pass first with the intermediate image rounded to uint8, then the vertical pass;
  * Pillow ``ImageEnhance`` Brightness / Contrast / Color = ``Image.blend(degenerate, image, factor)``
    (src/libImaging/Blend.c) with degenerate = black / the rounded mean of the L image / the L image, L =
    (19595 R + 38470 G + 7471 B + 0x8000) >> 16 (src/libImaging/Convert.c);
  * torchvision ``ToTensor`` (uint8 HWC -> float32 CHW / 255) and ``Normalize`` ((x - mean) / std in float32).

PINNED: tests/golden/make_golden_transforms.py runs the real PIL 12.2 / torchvision 0.26 of the authoring container on
seeded images and explicit augmentation parameters; tests/test_transforms_oracle.py requires this file to reproduce
those outputs bit for bit (and re-checks against the live libraries when they are importable).

Random parameter draws (crop box, jitter factors, flip) are NOT part of the arithmetic contract: they are explicit
inputs here, in the device kernel, and in the golden vectors.  ``random_resized_crop_params`` restates torchvision's
sampling logic on a ``torch.Generator`` for the host-side sampler.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = (0.485, 0.456, 0.406)      # data/datamgr.py:16
STD = (0.229, 0.224, 0.225)


# ---------------------------------------------------------------------------------------------- Pillow resample
def _bilinear(x):
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size, in0, in1, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter.  in0 / in1 are the box edges
    (C floats in Pillow; integers for every call the reference makes).  Returns (bounds [out,2] int, coeffs
    [out, ksize] int32) -- Python floats are IEEE doubles, the same arithmetic as the C code."""
    in0 = float(np.float32(in0))
    in1 = float(np.float32(in1))
    scale = filterscale = float(np.float32(in1) - np.float32(in0)) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One fixed-point pass along ``axis`` (1 = horizontal, 0 = vertical) of an HWC uint8 image."""
    src = img.astype(np.int64)
    n_out = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = n_out
    out = np.empty(shape, np.uint8)
    for o in range(n_out):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        k = kk[o, :n].astype(np.int64)
        if axis == 1:
            acc = (src[:, lo:lo + n, :] * k[None, :, None]).sum(1)
        else:
            acc = (src[lo:lo + n, :, :] * k[:, None, None]).sum(0)
        acc = (acc + (1 << (PRECISION_BITS - 1))) >> PRECISION_BITS
        acc = np.clip(acc, 0, 255).astype(np.uint8)
        if axis == 1:
            out[:, o, :] = acc
        else:
            out[o, :, :] = acc
    return out


def resize(img, out_w, out_h):
    """``Image.resize((out_w, out_h), BILINEAR)`` of an HWC uint8 array (the whole image is the box)."""
    h, w = img.shape[:2]
    if (w, h) == (out_w, out_h):
        return img.copy()
    bh, kh = precompute_coeffs(w, 0, w, out_w)
    bv, kv = precompute_coeffs(h, 0, h, out_h)
    # Pillow skips a pass whose size does not change; with the bilinear filter that pass is the identity anyway
    tmp = _pass(img, bh, kh, 1) if out_w != w else img
    return _pass(tmp, bv, kv, 0) if out_h != h else tmp


def resized_crop(img, top, left, h, w, out_size):
    """torchvision F.resized_crop on a PIL image: ``img.crop((left, top, left+w, top+h)).resize((S, S), BILINEAR)``."""
    return resize(img[top:top + h, left:left + w, :], out_size, out_size)


def center_crop(img, size):
    """torchvision CenterCrop(size) for an image at least ``size`` large in both directions."""
    h, w = img.shape[:2]
    top = int(round((h - size) / 2.0))
    left = int(round((w - size) / 2.0))
    return img[top:top + size, left:left + size, :]


# ---------------------------------------------------------------------------------------------- Pillow ImageEnhance
def to_L(img):
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(im1, im2, alpha):
    """Blend.c: out = im1 + alpha * (im2 - im1) in C float arithmetic, truncated to uint8 (clipped outside [0, 1])."""
    alpha = np.float32(alpha)
    if alpha == 0.0:
        return im1.copy()
    if alpha == 1.0:
        return im2.copy()
    a = im1.astype(np.int32)
    d = (im2.astype(np.int32) - a).astype(np.float32)
    t = a.astype(np.float32) + alpha * d          # float32 multiply, then float32 add (no fused multiply-add)
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    out = np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32)))
    return out.astype(np.uint8)


def brightness(img, f):
    return blend(np.zeros_like(img), img, f)


def contrast(img, f):
    L = to_L(img)
    mean = int(int(L.astype(np.int64).sum()) / L.size + 0.5)
    return blend(np.full_like(img, mean), img, f)


def color(img, f):
    L = to_L(img)
    return blend(np.repeat(L[..., None], 3, axis=2), img, f)


def image_jitter(img, factors):
    """data/additional_transforms.py:24-34 with the three factors given (dict order Brightness, Contrast, Color)."""
    out = brightness(img, factors[0])
    out = contrast(out, factors[1])
    out = color(out, factors[2])
    return out


# ---------------------------------------------------------------------------------------------- ToTensor / Normalize
def to_tensor_normalize(img, mean=MEAN, std=STD):
    x = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (x - m) / s


# ---------------------------------------------------------------------------------------------- the two pipelines
def transform_aug(img, crop, factors, flip, size, mean=MEAN, std=STD):
    """crop = (top, left, h, w).  data/datamgr.py:39 order."""
    top, left, h, w = crop
    out = resized_crop(img, top, left, h, w, size)
    if factors is not None:
        out = image_jitter(out, factors)
    if flip:
        out = out[:, ::-1, :]
    return to_tensor_normalize(np.ascontiguousarray(out), mean, std)


def transform_plain(img, size, mean=MEAN, std=STD):
    """data/datamgr.py:41: Scale([int(1.15 S)] * 2) -> CenterCrop(S) -> ToTensor -> Normalize."""
    big = int(size * 1.15)
    return to_tensor_normalize(np.ascontiguousarray(center_crop(resize(img, big, big), size)), mean, std)


# ---------------------------------------------------------------------------------------------- parameter sampling
def random_resized_crop_params(height, width, gen, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """torchvision RandomResizedCrop.get_params on a torch.Generator -> (top, left, h, w)."""
    import torch
    area = height * width
    log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1], generator=gen).item()
        aspect = math.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1], generator=gen).item())
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= width and 0 < h <= height:
            top = int(torch.randint(0, height - h + 1, (1,), generator=gen).item())
            left = int(torch.randint(0, width - w + 1, (1,), generator=gen).item())
            return top, left, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w
