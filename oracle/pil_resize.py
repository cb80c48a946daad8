"""ORACLE (test infrastructure): numpy restatement of Pillow's 8-bit bicubic resize (src/libImaging/Resample.c:
precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) and of the timm eval
transform built on it (Resize(shorter side) -> CenterCrop -> ToTensor -> Normalize), i.e. what the reference calls
`model.processor(PIL image)` (/root/reference/models/global_eval_models.py:333-334,
/root/reference/datasets/global_eval_torch_ds.py:20-21).  Pinned bit-exactly against Pillow itself in
tests/test_preprocess.py.  The same coefficient tables drive the CUDA kernel, so they are computed here in the same
double arithmetic as Pillow's C code."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """-> bounds [out,2] (xmin, count), coeffs int32 [out, ksize] in Q22 fixed point."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)], dtype=np.float64)
        ww = 0.0
        for v in w:          # same left-to-right double summation as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    q = np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))
    return bounds, np.trunc(q).astype(np.int32), ksize


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    in_size = img.shape[axis]
    bounds, kk, _ = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :cnt].astype(np.int64), src[xmin:xmin + cnt], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img uint8 [H, W, C]; horizontal pass first, then vertical, uint8 between the passes (Pillow's order)."""
    x = img
    if out_w != img.shape[1]:
        x = _resample_axis(x, out_w, 1)
    if out_h != img.shape[0]:
        x = _resample_axis(x, out_h, 0)
    return x


def resized_shape(h: int, w: int, size: int):
    """torchvision Resize(int): shorter side -> size, longer side int(size * long / short)."""
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


def eval_transform(img: np.ndarray, resize_to: int, crop: int, mean, std) -> np.ndarray:
    """uint8 [H, W, 3] -> float32 [3, crop, crop]; timm eval transform (resize_to = floor(crop / crop_pct))."""
    h, w = img.shape[:2]
    oh, ow = resized_shape(h, w, resize_to)
    r = resize_bicubic_u8(img, oh, ow)
    top, left = int(round((oh - crop) / 2.0)), int(round((ow - crop) / 2.0))
    c = r[top:top + crop, left:left + crop].astype(np.float32) / np.float32(255.0)
    c = (c - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(c.transpose(2, 0, 1))
