"""ORACLE (test infrastructure, not product code) -- PIL's 8-bit bicubic resample, restated.

The reference's `transform` (build-index.py:48 -> clip._transform -> torchvision
Resize(224, BICUBIC) + CenterCrop(224) on a PIL image) runs Pillow's ImagingResample: a
horizontal then a vertical pass, each with per-output-pixel windows of bicubic (a = -0.5)
weights, support 2 x max(scale, 1), normalised, rounded to 22-bit fixed point, accumulated
in int32 from 1 << 21 and shifted/clamped to uint8 between the passes.  This numpy version
is PINNED against Pillow itself (installed in this image): tests/test_resize.py asserts
bit-exact equality with `Image.resize(..., Image.BICUBIC)` + crop over many sizes.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size: int, out_size: int):
    """Per output pixel: (xmin, count) bounds and int32 fixed-point weights [out_size, ksize]."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resize_size(w: int, h: int, n_px: int = 224):
    """torchvision Resize(int): shorter side -> n_px, longer side truncated."""
    if w <= h:
        return n_px, int(n_px * h / w)
    return int(n_px * w / h), n_px


def crop_origin(nw: int, nh: int, n_px: int = 224):
    return int(round((nw - n_px) / 2.0)), int(round((nh - n_px) / 2.0))


def _pass(img: np.ndarray, bounds, kk, axis: int) -> np.ndarray:
    """One resample pass along `axis` (1 = horizontal, 0 = vertical) of an [H, W, C] uint8 image."""
    src = img.astype(np.int64)
    out_size = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.empty(shape, dtype=np.uint8)
    for o in range(out_size):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        k = kk[o, :n].astype(np.int64)
        if axis == 1:
            acc = (src[:, lo:lo + n, :] * k[None, :, None]).sum(axis=1)
        else:
            acc = (src[lo:lo + n, :, :] * k[:, None, None]).sum(axis=0)
        acc = (acc + (1 << (PRECISION_BITS - 1))) >> PRECISION_BITS
        res = np.clip(acc, 0, 255).astype(np.uint8)
        if axis == 1:
            out[:, o, :] = res
        else:
            out[o, :, :] = res
    return out


def resize_center_crop(img: np.ndarray, n_px: int = 224) -> np.ndarray:
    """[H, W, 3] uint8 -> [n_px, n_px, 3] uint8, bit-identical to
    PIL.Image.resize((nw, nh), BICUBIC) + center crop (for RGB images)."""
    h, w = img.shape[:2]
    if (w, h) == (n_px, n_px):
        return img.copy()
    nw, nh = resize_size(w, h, n_px)
    left, top = crop_origin(nw, nh, n_px)
    cur = img
    if nw != w:
        bx, kx = coeffs(w, nw)
        cur = _pass(cur, bx, kx, axis=1)
    if nh != h:
        by, ky = coeffs(h, nh)
        cur = _pass(cur, by, ky, axis=0)
    return cur[top:top + n_px, left:left + n_px].copy()
