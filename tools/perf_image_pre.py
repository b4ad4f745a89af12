"""Throughput of the image pipeline on the GPU box (dev tool): B200ViltFeatureExtractor vs the reference arithmetic on the host
(Pillow resize + numpy normalise + pad = oracle.image_oracle would be a port; here Pillow itself does the resize, as in the
reference). 64 COCO-sized uint8 images per batch, host arrays in, device tensors out (H2D copy inside the timed region)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200.image_processing import B200ViltFeatureExtractor, target_size  # noqa: E402


def main():
    B = int(os.environ.get("B", 64))
    rng = np.random.default_rng(0)
    shapes = [(480, 640), (640, 480), (427, 640), (500, 375)]
    imgs = [rng.integers(0, 256, shapes[i % 4] + (3,), dtype=np.uint8) for i in range(B)]
    fe = B200ViltFeatureExtractor()
    for _ in range(3):
        out = fe(imgs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        out = fe(imgs)
    torch.cuda.synchronize()
    gpu = (time.perf_counter() - t0) / n
    # device part alone (CUDA events around one call's kernels would need the plan split; the host share is the rest)
    from PIL import Image
    pil = [Image.fromarray(im) for im in imgs[:16]]
    t0 = time.perf_counter()
    for im in pil:
        oh, ow = target_size(im.size[1], im.size[0])
        x = np.asarray(im.resize((ow, oh), resample=Image.BICUBIC)).astype(np.float32) / 255.0
        x = (x.transpose(2, 0, 1) - 0.5) / 0.5
    cpu = (time.perf_counter() - t0) / 16
    print(f"image pipeline: B={B} {gpu * 1e3:.2f} ms per batch = {B / gpu:.0f} images/s end to end (host arrays -> device tensors); "
          f"Pillow + numpy on one host core: {cpu * 1e3:.2f} ms per image = {1 / cpu:.0f} images/s; output {tuple(out['pixel_values'].shape)}")


if __name__ == "__main__":
    main()
