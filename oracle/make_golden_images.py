"""Golden vectors of the UNMODIFIED reference image path: ViltFeatureExtractor.__call__
(adapter-transformers/src/transformers/models/vilt/feature_extraction_vilt.py:175-292, the class behind
ViltEncoderWrapper.process_inputs, src/modeling/vilt.py:83-96) run here on CPU, with Pillow doing the bicubic resize.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_images        # writes tests/golden/image_pre_*.npz

TEST INFRASTRUCTURE: build container only (the GPU box has no /root/reference). Inputs are synthetic uint8 images from a
seeded generator (smooth gradients + blocks + noise, so that both the ringing of the bicubic kernel and the 8-bit clipping
show up); the fixtures hold inputs and outputs. `size` is 64 / 96 for the small fixtures (the extractor's own parameter:
shorter edge -> size, longer edge <= int(1333 / 800 * size)) and the default 384 for one smooth pair, to keep the files small.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def synth_image(rng, h, w, smooth=False):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    f = 12.0 if smooth else 1.0                # the default-size fixture is low-frequency so that it compresses
    img = np.stack([127.5 + 127.5 * np.sin(xx / (f * (3.0 + c)) + yy / (f * 7.0)) * np.cos(yy / (f * (5.0 + 2 * c))) for c in range(3)], -1)
    if not smooth:
        img += rng.normal(0, 40, img.shape)
        for _ in range(6):                       # saturated blocks: overshoot of the cubic kernel gets clipped
            y0, x0 = rng.integers(0, max(1, h - 4)), rng.integers(0, max(1, w - 4))
            img[y0:y0 + rng.integers(2, 12), x0:x0 + rng.integers(2, 12)] = rng.choice([0.0, 255.0])
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def main():
    ref_shim.install()
    from PIL import Image
    import PIL
    from transformers import ViltFeatureExtractor
    rng = np.random.default_rng(20260117)
    cases = {
        "image_pre_size64": (64, [(48, 64), (97, 61), (40, 100), (64, 64), (33, 35)], False),
        "image_pre_size96": (96, [(200, 150), (96, 160), (77, 301)], False),
        "image_pre_size384": (384, [(120, 160), (400, 520)], True),
    }
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, (size, shapes, smooth) in cases.items():
        imgs = [synth_image(rng, h, w, smooth) for h, w in shapes]
        fe = ViltFeatureExtractor(size=size)
        out = fe([Image.fromarray(im) for im in imgs], return_tensors="np")
        pv, pm = out["pixel_values"].astype(np.float32), out["pixel_mask"].astype(np.int64)
        # compact, lossless encoding of the float32 output: every real pixel is one of <= 256 float32 values (a normalised
        # uint8); the fixture stores the index per pixel and the table of the REFERENCE's own float32 values
        k = np.rint((pv.astype(np.float64) * 0.5 + 0.5) * 255.0).astype(np.int64)
        real = np.broadcast_to(pm[:, None] == 1, pv.shape)
        assert k[real].min() >= 0 and k[real].max() <= 255 and np.all(pv[~real] == 0.0)
        lut = np.full(256, np.nan, np.float32)
        for v in np.unique(k[real]):
            vals = np.unique(pv[real & (k == v)])
            assert vals.size == 1
            lut[v] = vals[0]
        k8 = np.where(real, k, 0).astype(np.uint8)
        assert np.array_equal(np.where(real, lut[k8], np.float32(0.0)), pv)
        data = {"size": np.int64(size), "n": np.int64(len(imgs)), "pillow": np.array(PIL.__version__),
                "pixel_index": k8, "lut": lut, "pixel_mask": pm}
        for i, im in enumerate(imgs):
            data[f"image_{i}"] = im
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **data)
        print(name, pv.shape, pm.sum(axis=(1, 2)))


if __name__ == "__main__":
    main()
