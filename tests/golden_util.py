"""Shared helpers for replaying tests/golden/*.npz (written by oracle/make_golden.py from the
unmodified reference)."""
import os

import numpy as np
import torch

from oracle.vilt_oracle import ViltDims, synth_batch, synth_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_TASKS = ["vqa", "nlvr2", "snli-ve", "vcr"]
TINY = ViltDims(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                image_size=32, patch_size=16, vocab_size=200, max_position_embeddings=8)
TINY_HW = (48, 64)
TINY_T = 8
BASE = ViltDims()
BASE_HW = (448, 448)
GRAD_SAMPLES = 2048


def load(tag):
    return np.load(os.path.join(GOLDEN_DIR, tag + ".npz"), allow_pickle=False)


def fixture_scales(g):
    """layer_scale / head_scale the fixture's weights were generated with (the VCR fixtures: see
    oracle.vilt_oracle.synth_state_dict); 1.0 for fixtures written before the knobs existed."""
    return dict(layer_scale=float(g["layer_scale"]) if "layer_scale" in g.files else 1.0,
                head_scale=float(g["head_scale"]) if "head_scale" in g.files else 1.0)


def grad_sample_index(numel):
    if numel <= GRAD_SAMPLES:
        return np.arange(numel)
    return (np.arange(GRAD_SAMPLES, dtype=np.int64) * (numel - 1)) // (GRAD_SAMPLES - 1)


def regen_batch(g, task, dims, T, hw, B, seed, masked):
    """Regenerate the synthetic batch and check it against what the reference was fed. Fixtures with
    `image_sizes` are padded batches (images of different sizes, pixel_mask zeros)."""
    batch = synth_batch(task, B, dims, T=T, image_hw=hw, seed=seed, masked=masked)
    if "image_sizes" in g.files:
        from oracle.vilt_oracle import pad_batch_images
        batch = pad_batch_images(batch, [tuple(x) for x in g["image_sizes"].tolist()])
    for k in ("input_ids", "attention_mask", "token_type_ids", "target"):
        assert np.array_equal(g["in_" + k], batch[k].numpy()), k
    px = batch["pixel_values"].double()
    chk = np.array([px.sum().item(), px.abs().sum().item(), (px * px).sum().item()])
    assert np.allclose(chk, g["in_pixel_checksum"], rtol=1e-12), "synthetic pixels differ from the golden run"
    return batch


def compare_grads(g, grads, rtol_norm, tol_elem, names=None):
    """grads: {name: tensor}. Checks every gradient the reference produced: norm, and full tensor
    or strided sample, relative to the reference gradient's max-abs."""
    checked = 0
    worst = (0.0, None)
    # gradients that are analytically zero (e.g. the key bias: softmax is shift-invariant) are pure
    # rounding noise in both implementations: compare them against the overall gradient scale
    global_scale = max(float(g[k]) for k in g.files if k.startswith("gnorm/"))
    for key in g.files:
        if not key.startswith("gnorm/"):
            continue
        name = key[len("gnorm/"):]
        if names is not None and name not in names:
            continue
        assert name in grads and grads[name] is not None, f"missing gradient for {name}"
        got = grads[name].detach().float().cpu()
        ref_norm = float(g[key])
        if ref_norm < 1e-6 * global_scale:
            assert got.norm().item() < 1e-5 * global_scale, (name, got.norm().item(), ref_norm)
            checked += 1
            continue
        assert abs(got.norm().item() - ref_norm) <= rtol_norm * max(ref_norm, 1e-12), (
            name, got.norm().item(), ref_norm)
        if "grad/" + name in g.files:
            ref = torch.from_numpy(g["grad/" + name])
            got_c = got.reshape(ref.shape)
        else:
            ref = torch.from_numpy(g["gsample/" + name])
            got_c = got.flatten()[torch.from_numpy(grad_sample_index(got.numel()))]
        scale = max(ref.abs().max().item(), 1e-12)
        err = (got_c - ref).abs().max().item() / scale
        if err > worst[0]:
            worst = (err, name)
        assert err <= tol_elem, (name, err)
        checked += 1
    assert checked > 0
    return worst
