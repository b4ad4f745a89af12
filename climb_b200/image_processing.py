"""B200 image pipeline: drop-in for the image half of `ViltProcessor` / `ViltFeatureExtractor.__call__`
(adapter-transformers/src/transformers/models/vilt/feature_extraction_vilt.py:175-292), which CLiMB calls per batch from
`ViltEncoderWrapper.process_inputs` (src/modeling/vilt.py:83-96) on the CPU with PIL.

    fe = B200ViltFeatureExtractor(size=384)                  # same parameters as ViltFeatureExtractor
    out = fe(images)                                         # list of PIL images / [H, W, 3] uint8 arrays / [3, H, W] uint8 tensors
    out["pixel_values"]  # cuda float32 [B, 3, Hmax, Wmax], out["pixel_mask"]  # cuda int64 [B, Hmax, Wmax]

Same results as the reference, bit for bit: the target size rule (:109-127), Pillow's 8-bit BICUBIC resampling, x / 255 and
(x - mean) / std in float32, zero padding to the batch maximum, the 0/1 pixel mask. What runs where:
  host    the size rule, Pillow's coefficient tables (a few hundred int32 per image axis: `resample_tables`, cached per
          (input size, output size)), packing the uint8 images into ONE pinned staging buffer, one async H2D copy
  device  both resampling passes, normalisation, padding, mask: two launches per batch (csrc/image_pre.cu)
The uint8 images (0.9 MB for 480 x 640) cross PCIe instead of the float32 result (2.4 MB at 384 x 512), and the CPU work per
image drops from a full PIL resize (~5 ms) to a memcpy.

There is no CPU fallback: without the CUDA library this module raises on import, with CPU tensors it raises on use.
"""
from __future__ import annotations

import ctypes
import functools
import math

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2                    # Pillow, src/libImaging/Resample.c


def target_size(h, w, shorter=384, size_divisor=32):
    """ViltFeatureExtractor._resize (feature_extraction_vilt.py:109-127): shorter edge -> `shorter`, longer edge limited to
    int(1333 / 800 * shorter), both floored to a multiple of `size_divisor`. Returns (new_h, new_w)."""
    longer = int((1333 / 800) * shorter)
    scale = shorter / min(w, h)
    if h < w:
        newh, neww = shorter, scale * w
    else:
        newh, neww = scale * h, shorter
    if max(newh, neww) > longer:
        scale = longer / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    newh, neww = int(newh + 0.5), int(neww + 0.5)
    return newh // size_divisor * size_divisor, neww // size_divisor * size_divisor


@functools.lru_cache(maxsize=4096)
def resample_tables(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for BICUBIC over the full axis, vectorised over the output
    positions with the operation order of the C code (the window sum runs left to right: np.cumsum, not np.sum).
    Returns (bounds [out, 2] int32 = (first input index, tap count), coeffs [out, ksize] int32, ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)            # (int) truncates; negative values clamp to 0 anyway
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    t = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * ss)
    a = -0.5
    w = np.where(t < 1.0, ((a + 2.0) * t - (a + 3.0)) * t * t + 1, np.where(t < 2.0, (((t - 5) * t + 8) * t - 4) * a, 0.0))
    w = np.where(x < xmax[:, None], w, 0.0)
    ww = np.cumsum(w, axis=1)[:, -1:]
    w = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    kk = np.where(w < 0, np.trunc(-0.5 + w * (1 << PRECISION_BITS)), np.trunc(0.5 + w * (1 << PRECISION_BITS))).astype(np.int32)
    bounds = np.stack([xmin, xmax], 1).astype(np.int32)
    return bounds, kk, ksize


_POOL = None


def _pack_pool():
    global _POOL
    if _POOL is None:
        import concurrent.futures
        import os
        _POOL = concurrent.futures.ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 4) // 2)), thread_name_prefix="climb_b200_pack")
    return _POOL


def _as_hwc_uint8(image):
    """PIL image / [H, W, 3] uint8 array / [3, H, W] uint8 tensor or array -> contiguous [H, W, 3] uint8 numpy array."""
    if isinstance(image, torch.Tensor):
        image = image.detach().cpu().numpy()
    if not isinstance(image, np.ndarray):
        if getattr(image, "mode", "RGB") != "RGB":
            image = image.convert("RGB")              # the reference's datasets do the same before the processor
        image = np.asarray(image)
    if image.ndim != 3:
        raise ValueError(f"expected an RGB image, got an array of shape {image.shape}")
    if image.shape[2] != 3 and image.shape[0] == 3:
        image = image.transpose(1, 2, 0)
    if image.shape[2] != 3 or image.dtype != np.uint8:
        raise ValueError(f"expected uint8 RGB, got dtype {image.dtype} shape {image.shape}")
    return np.ascontiguousarray(image)


class B200ViltFeatureExtractor:
    """Same constructor arguments and output keys as ViltFeatureExtractor (do_resize / do_normalize are always on: the
    reference's defaults, the only configuration CLiMB uses); `resample` must be BICUBIC (3)."""

    model_input_names = ["pixel_values", "pixel_mask"]

    def __init__(self, size=384, size_divisor=32, resample=3, image_mean=None, image_std=None, device="cuda"):
        if resample != 3:
            raise ValueError("B200ViltFeatureExtractor implements Pillow's BICUBIC (3) only")
        self.size, self.size_divisor = size, size_divisor
        self.image_mean = list(image_mean) if image_mean is not None else [0.5, 0.5, 0.5]
        self.image_std = list(image_std) if image_std is not None else [0.5, 0.5, 0.5]
        self.device = torch.device(device)
        self._stage = None            # pinned staging buffer, grown on demand
        self._copied = None           # event behind the last H2D copy out of it

    def __deepcopy__(self, memo):
        # trainers deep-copy the whole learner for best-model tracking (train_vqa.py:210): parameters only, no staging buffer / event
        return B200ViltFeatureExtractor(self.size, self.size_divisor, 3, self.image_mean, self.image_std, self.device)

    def plan(self, shapes):
        """Host-side plan of a batch: per image the descriptor fields, the concatenated int32 tables, the canvas size."""
        tables, descs, t_off, s_off, tmp_off, max_tmp = [], [], 0, 0, 0, 0
        hp = wp = 0
        for (h, w) in shapes:
            oh, ow = target_size(h, w, self.size, self.size_divisor)
            if oh <= 0 or ow <= 0:
                raise ValueError("height and width must be > 0")        # Pillow's message for the same degenerate aspect ratios
            bh, kh, ksh = resample_tables(w, ow)
            bv, kv, ksv = resample_tables(h, oh)
            offs = []
            for arr in (bh, kh, bv, kv):
                offs.append(t_off)
                tables.append(arr.reshape(-1))
                t_off += arr.size
            descs.append((s_off, tmp_off, h, w, oh, ow, ksh, ksv, offs[0], offs[1], offs[2], offs[3]))
            s_off += h * w * 3
            tmp_off += h * ow * 3
            max_tmp = max(max_tmp, h * ow)
            hp, wp = max(hp, oh), max(wp, ow)
        return descs, np.concatenate(tables), s_off, tmp_off, max_tmp, hp, wp

    def __call__(self, images, return_tensors="pt"):
        if self.device.type != "cuda":
            raise RuntimeError("B200ViltFeatureExtractor runs on a CUDA device only (no CPU fallback)")
        if not isinstance(images, (list, tuple)):
            images = [images]
        imgs = [_as_hwc_uint8(im) for im in images]
        descs, tables, src_bytes, tmp_bytes, max_tmp, hp, wp = self.plan([im.shape[:2] for im in imgs])
        B = len(imgs)
        desc_arr = (_lib.ImageDescC * B)(*[_lib.ImageDescC(*d) for d in descs])
        desc_bytes, table_bytes = ctypes.sizeof(desc_arr), tables.nbytes
        # one pinned staging buffer: [descriptors | tables | pixels], one H2D copy
        o_tab = (desc_bytes + 15) // 16 * 16
        o_src = (o_tab + table_bytes + 15) // 16 * 16
        total = o_src + src_bytes
        if self._copied is not None:
            self._copied.synchronize()                # the previous batch's copy must have left the staging buffer
        if self._stage is None or self._stage.numel() < total:
            self._stage = torch.empty(int(total * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
        stage = self._stage.numpy()
        stage[:desc_bytes] = np.frombuffer(desc_arr, dtype=np.uint8)
        stage[o_tab:o_tab + table_bytes] = tables.view(np.uint8)
        # packing the pixels into the staging buffer is the host's whole share of the work (0.9 MB per COCO image): a few
        # threads do it, numpy's copy releases the GIL
        jobs, off = [], o_src
        for im in imgs:
            jobs.append((off, im))
            off += im.size

        def put(job):
            stage[job[0]:job[0] + job[1].size] = job[1].reshape(-1)

        if src_bytes >= (4 << 20) and len(jobs) >= 4:
            list(_pack_pool().map(put, jobs))
        else:
            for job in jobs:
                put(job)
        dev = self._stage[:total].to(self.device, non_blocking=True)
        if self._copied is None:
            self._copied = torch.cuda.Event()
        self._copied.record()
        tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=self.device)
        pv = torch.empty(B, 3, hp, wp, dtype=torch.float32, device=self.device)
        pm = torch.empty(B, hp, wp, dtype=torch.int64, device=self.device)
        mean = (ctypes.c_float * 3)(*self.image_mean)
        std = (ctypes.c_float * 3)(*self.image_std)
        base = dev.data_ptr()
        _lib.check(_lib.climb_image_preprocess(base + o_src, _lib.ptr(tmp), base, base + o_tab, B, max_tmp, _lib.ptr(pv), _lib.ptr(pm),
                                               hp, wp, mean, std, _lib.stream()))
        return {"pixel_values": pv, "pixel_mask": pm}
