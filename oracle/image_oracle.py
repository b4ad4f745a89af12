"""CPU restatement of the image side of `ViltEncoderWrapper.process_inputs` (src/modeling/vilt.py:83-96 ->
ViltProcessor -> ViltFeatureExtractor): resize, normalise, pad, pixel mask.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product path (climb_b200/image_processing.py + csrc/image_pre.cu) never does.

What is restated, and from where:
  * target size          ViltFeatureExtractor._resize, feature_extraction_vilt.py:90-127 (shorter edge -> 384, longer edge
                         <= int(1333 / 800 * 384) = 640, both floored to a multiple of 32) and __call__ :253-263
  * bicubic resampling   a THIRD-PARTY dependency of the reference, absent from /root/reference: Pillow's
                         `Image.resize(size, resample=Image.BICUBIC)` (R/requirements.txt does not pin Pillow; this image has
                         12.2.0, whose src/libImaging/Resample.c has carried the same algorithm since 3.4 / 5.x):
                         precompute_coeffs -> normalize_coeffs_8bpc (PRECISION_BITS = 22) -> ImagingResampleHorizontal_8bpc
                         -> ImagingResampleVertical_8bpc, the horizontal pass first, uint8 between the passes. The published
                         algorithm is restated below in integer arithmetic; tests/test_image_pre.py pins it bit-exactly to
                         Pillow itself and to golden vectors of the UNMODIFIED ViltFeatureExtractor (oracle/make_golden_images.py).
  * normalise            FeatureExtractionMixin.to_numpy_array (x.astype(float32) / 255.0, channel first) and .normalize
                         ((x - mean) / std in float32, mean = std = 0.5), feature_extraction_utils.py
  * pad + pixel mask     feature_extraction_vilt.py:265-283 (zeros up to the batch maximum, mask 1 on real pixels, int64)
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c


def target_size(h: int, w: int, shorter: int = 384, size_divisor: int = 32):
    """(new_h, new_w) of ViltFeatureExtractor._resize (feature_extraction_vilt.py:109-127); longer = int(1333 / 800 * size)."""
    longer = int((1333 / 800) * shorter)
    scale = shorter / min(w, h)
    if h < w:
        newh, neww = shorter, scale * w
    else:
        newh, neww = scale * h, shorter
    if max(newh, neww) > longer:
        scale = longer / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    newh, neww = int(newh + 0.5), int(neww + 0.5)
    return newh // size_divisor * size_divisor, neww // size_divisor * size_divisor


def _bicubic(x: float) -> float:
    a = -0.5                          # Resample.c: bicubic_filter
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full box [0, in_size): (bounds [out, 2] int32 =
    (first input index, tap count), coeffs [out, ksize] int32 in 2^-22 units)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
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
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit pass along `axis` (0 = vertical, 1 = horizontal) of an [H, W, C] uint8 image."""
    bounds, kk = resample_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)                  # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow's Image.resize((out_w, out_h), BICUBIC) on an [H, W, C] uint8 array: the horizontal pass first, a pass is
    skipped when the size along it does not change (Resample.c ImagingResampleInner)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    if img.shape[1] != out_w:
        img = _resample_axis(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _resample_axis(img, out_h, 0)
    return img


def normalize_chw(img_u8: np.ndarray, mean: float = 0.5, std: float = 0.5) -> np.ndarray:
    """to_numpy_array(rescale, channel_first) + normalize of feature_extraction_utils.py, float32 throughout."""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = x.transpose(2, 0, 1)
    return (x - np.float32(mean)) / np.float32(std)


def feature_extract(images, shorter: int = 384, size_divisor: int = 32):
    """ViltFeatureExtractor.__call__ (feature_extraction_vilt.py:253-292) on a list of [H, W, 3] uint8 arrays ->
    (pixel_values [B, 3, Hmax, Wmax] float32, pixel_mask [B, Hmax, Wmax] int64)."""
    out = []
    for im in images:
        nh, nw = target_size(im.shape[0], im.shape[1], shorter, size_divisor)
        out.append(normalize_chw(resize_bicubic_u8(im, nh, nw)))
    hm, wm = max(o.shape[1] for o in out), max(o.shape[2] for o in out)
    pv = np.zeros((len(out), 3, hm, wm), np.float32)
    pm = np.zeros((len(out), hm, wm), np.int64)
    for i, o in enumerate(out):
        pv[i, :, :o.shape[1], :o.shape[2]] = o
        pm[i, :o.shape[1], :o.shape[2]] = 1
    return pv, pm
