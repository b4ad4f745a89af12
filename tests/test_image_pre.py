"""Image side of the input pipeline: oracle/image_oracle.py against the golden vectors of the
UNMODIFIED ViltFeatureExtractor (oracle/make_golden_images.py) and against Pillow itself; the host planning code against
the oracle; the CUDA path (climb_b200/image_processing.py + csrc/image_pre.cu) against both. Everything is integer / exactly
rounded float32 work: the bar is bit-exact."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from oracle import image_oracle as io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "image_pre_*.npz")))


def _load(path):
    d = np.load(path)
    imgs = [d[f"image_{i}"] for i in range(int(d["n"]))]
    pm = d["pixel_mask"]
    real = np.broadcast_to(pm[:, None] == 1, d["pixel_index"].shape)
    pv = np.where(real, d["lut"][d["pixel_index"]], np.float32(0.0)).astype(np.float32)     # the reference's own float32 values
    return int(d["size"]), imgs, pv, pm


def _rand_images(seed, shapes):
    rng = np.random.default_rng(seed)
    out = []
    for h, w in shapes:
        im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        im[:: max(2, h // 7)] = 255          # saturated lines: the cubic kernel's overshoot has to be clipped
        im[:, :: max(2, w // 5)] = 0
        out.append(im)
    return out


# ------------------------------------------------------------------------------------------------ CPU: the oracle is pinned
def test_golden_fixtures_present():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_feature_extractor(path):
    size, imgs, pv, pm = _load(path)
    got_pv, got_pm = io.feature_extract(imgs, shorter=size)
    assert got_pv.dtype == np.float32 and got_pm.dtype == np.int64
    assert np.array_equal(got_pm, pm)
    assert np.array_equal(got_pv, pv)


def test_oracle_resampling_is_pillow_bit_for_bit():
    Image = pytest.importorskip("PIL.Image")
    cases = [(48, 64, 32, 32), (97, 61, 96, 64), (37, 53, 96, 64), (100, 100, 100, 160), (50, 70, 20, 30), (7, 9, 32, 32),
             (200, 300, 64, 96), (64, 96, 64, 96)]
    for (h, w, oh, ow), im in zip(cases, _rand_images(5, [c[:2] for c in cases])):
        ref = np.asarray(Image.fromarray(im).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(io.resize_bicubic_u8(im, oh, ow), ref), (h, w, oh, ow)


def test_target_size_rule():
    # feature_extraction_vilt.py:109-127 at the default size 384 (longer edge <= 639, multiples of 32)
    assert io.target_size(480, 640) == (384, 512)
    assert io.target_size(640, 480) == (512, 384)
    assert io.target_size(333, 500) == (384, 576)
    assert io.target_size(300, 1000) == (192, 608)      # longer = int(1333 / 800 * 384) = 639, floored to a multiple of 32
    assert io.target_size(448, 448) == (384, 384)


# ------------------------------------------------------------------------------------------------ CPU: host logic of the product
def test_host_planning_matches_oracle():
    from climb_b200 import _lib, image_processing as ip
    assert ctypes.sizeof(_lib.ImageDescC) == 2 * 8 + 6 * 4 + 4 * 8
    for h, w in [(480, 640), (640, 480), (333, 500), (97, 61), (40, 100), (1200, 1600), (384, 384)]:
        assert ip.target_size(h, w) == io.target_size(h, w)
        assert ip.target_size(h, w, 64) == io.target_size(h, w, 64)
    for i, o in [(640, 512), (480, 384), (500, 576), (53, 64), (37, 96), (70, 30), (100, 100), (1024, 512), (9, 32), (4000, 640)]:
        b, k = io.resample_coeffs(i, o)
        b2, k2, ks = ip.resample_tables(i, o)
        assert ks == k.shape[1] and np.array_equal(b, b2) and np.array_equal(k, k2), (i, o)
    fe = ip.B200ViltFeatureExtractor(size=64, device="cuda")
    descs, tables, src_bytes, tmp_bytes, max_tmp, hp, wp = fe.plan([(48, 64), (97, 61)])
    assert (hp, wp) == (96, 64) and src_bytes == (48 * 64 + 97 * 61) * 3 and tables.dtype == np.int32
    with pytest.raises(ValueError):
        fe.plan([(40, 200)])                           # the reference fails on this aspect ratio too (height 0)
    with pytest.raises(ValueError):
        ip.B200ViltFeatureExtractor(resample=2)
    with pytest.raises(RuntimeError):
        ip.B200ViltFeatureExtractor(device="cpu")([np.zeros((32, 32, 3), np.uint8)])


# ------------------------------------------------------------------------------------------------ GPU: the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_reference_feature_extractor(path):
    from climb_b200.image_processing import B200ViltFeatureExtractor
    size, imgs, pv, pm = _load(path)
    out = B200ViltFeatureExtractor(size=size)(imgs)
    assert out["pixel_values"].dtype == torch.float32 and out["pixel_mask"].dtype == torch.int64
    assert torch.equal(out["pixel_mask"].cpu(), torch.from_numpy(pm))
    assert torch.equal(out["pixel_values"].cpu(), torch.from_numpy(pv))


@pytest.mark.gpu
def test_cuda_matches_oracle_on_mixed_batches():
    from climb_b200.image_processing import B200ViltFeatureExtractor
    Image = pytest.importorskip("PIL.Image")
    # default size: COCO-like landscape / portrait, upscaling, a 4x downscale, an axis that keeps its size, tiny inputs
    shapes = [(480, 640), (640, 480), (120, 160), (1536, 2048), (384, 700), (33, 35), (427, 640)]
    imgs = _rand_images(11, shapes)
    fe = B200ViltFeatureExtractor()
    ref_pv, ref_pm = io.feature_extract(imgs)
    forms = [imgs,                                                       # [H, W, 3] arrays
             [Image.fromarray(im) for im in imgs],                       # PIL images (what CLiMB's datasets hand over)
             [torch.from_numpy(im).permute(2, 0, 1) for im in imgs]]     # [3, H, W] tensors
    for batch in forms:
        out = fe(batch)
        assert torch.equal(out["pixel_mask"].cpu(), torch.from_numpy(ref_pm))
        assert torch.equal(out["pixel_values"].cpu(), torch.from_numpy(ref_pv))
    # a second, smaller batch through the same (reused) staging buffer; single image without a list
    one = fe(imgs[2])
    r1, m1 = io.feature_extract(imgs[2:3])
    assert torch.equal(one["pixel_values"].cpu(), torch.from_numpy(r1)) and torch.equal(one["pixel_mask"].cpu(), torch.from_numpy(m1))


@pytest.mark.gpu
def test_cuda_custom_size_and_statistics():
    from climb_b200.image_processing import B200ViltFeatureExtractor
    imgs = _rand_images(3, [(97, 61), (64, 64), (200, 150)])
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    out = B200ViltFeatureExtractor(size=96, image_mean=mean, image_std=std)(imgs)
    res = [io.resize_bicubic_u8(im, *io.target_size(im.shape[0], im.shape[1], 96)) for im in imgs]
    hm, wm = max(r.shape[0] for r in res), max(r.shape[1] for r in res)
    ref = np.zeros((3, 3, hm, wm), np.float32)
    for i, r in enumerate(res):
        x = r.astype(np.float32) / np.float32(255.0)
        ref[i, :, :r.shape[0], :r.shape[1]] = ((x - np.array(mean, np.float32)) / np.array(std, np.float32)).transpose(2, 0, 1)
    assert torch.equal(out["pixel_values"].cpu(), torch.from_numpy(ref))


@pytest.mark.gpu
def test_process_inputs_runs_the_image_half_on_the_gpu():
    """B200ViltEncoderWrapper.process_inputs (= ViltEncoderWrapper.process_inputs, src/modeling/vilt.py:83-96) with a
    ViltProcessor-shaped object: the tokenizer's tensors pass through, pixel_values / pixel_mask equal the reference
    extractor's (through the oracle), and the encodings feed the encoder."""
    import types
    from climb_b200.modeling import B200ViltConfig, B200ViltEncoderWrapper, B200ViltModel
    cfg = B200ViltConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=64,
                         patch_size=32, vocab_size=200, max_position_embeddings=8)

    def tokenizer(text, max_length, padding, truncation, return_tensors):
        n = len(text)
        ids = torch.arange(1, 1 + n * 6).view(n, 6) % 199 + 1
        return {"input_ids": ids, "token_type_ids": torch.zeros_like(ids), "attention_mask": torch.ones_like(ids)}

    fe = types.SimpleNamespace(size=64, size_divisor=32, image_mean=[0.5, 0.5, 0.5], image_std=[0.5, 0.5, 0.5], do_resize=True,
                               do_normalize=True, resample=3)
    proc = types.SimpleNamespace(tokenizer=tokenizer, feature_extractor=fe)
    dev = torch.device("cuda")
    enc_wrapper = B200ViltEncoderWrapper(proc, B200ViltModel(cfg), dev).to(dev)
    imgs = _rand_images(21, [(48, 64), (97, 61), (64, 64)])
    enc = enc_wrapper.process_inputs(imgs, ["a b c", "d e", "f"])
    ref_pv, ref_pm = io.feature_extract(imgs, shorter=64)
    assert torch.equal(enc["pixel_values"].cpu(), torch.from_numpy(ref_pv)) and torch.equal(enc["pixel_mask"].cpu(), torch.from_numpy(ref_pm))
    assert enc["input_ids"].device.type == "cuda" and enc["input_ids"].shape == (3, 6)
    pooled = enc_wrapper(**enc)
    assert pooled.shape == (3, 128) and torch.isfinite(pooled).all()
