"""CPU: the oracle restatement (oracle/vilt_oracle.py) against the golden vectors produced by the
unmodified reference (oracle/make_golden.py). fp32 vs fp32 on the same machine: the only expected
differences are summation order and the reference's random patch permutation (~1e-6)."""
import numpy as np
import pytest
import torch

from oracle import vilt_oracle as vo
from tests.golden_util import (ALL_TASKS, BASE, BASE_HW, TINY, TINY_HW, TINY_T, compare_grads, fixture_scales, load,
                               regen_batch)


def _oracle_step(sd, dims, task, batch, adapter=None):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pooled, logits = vo.learner_forward(params, dims, task, batch, adapter=adapter)
    loss = vo.task_loss(task, logits, batch["target"])
    loss.backward()
    return pooled, logits, loss, {k: v.grad for k, v in params.items()}


@pytest.mark.parametrize("task", ALL_TASKS)
def test_tiny_tasks_match_reference(task):
    g = load(f"tiny_{task}")
    seed = int(g["seed"])
    batch = regen_batch(g, task, TINY, TINY_T, TINY_HW, 3, seed, True)
    assert np.array_equal(g["in_pixel_values"], batch["pixel_values"].numpy())
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=seed, **fixture_scales(g))
    pooled, logits, loss, grads = _oracle_step(sd, TINY, task, batch)
    assert np.allclose(pooled.detach().numpy(), g["pooled"], atol=2e-6)
    assert np.allclose(logits.detach().numpy(), g["logits"], atol=5e-6)
    assert abs(loss.item() - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    if task == "vcr":
        assert abs(float(g["loss"]) - np.log(4.0)) > 1e-2          # the multi-choice fixture is not the degenerate ln 4 one
    worst = compare_grads(g, grads, rtol_norm=1e-4, tol_elem=2e-4)
    print("worst grad", worst)


@pytest.mark.parametrize("tag,kind,task,rf", [("tiny_adapter_houlsby_nlvr2", "houlsby", "nlvr2", 4),
                                               ("tiny_adapter_pfeiffer_vqa", "pfeiffer", "vqa", 2)])
def test_tiny_adapters_match_reference(tag, kind, task, rf):
    g = load(tag)
    seed = int(g["seed"])
    batch = regen_batch(g, task, TINY, TINY_T, TINY_HW, 3, seed, True)
    sites = ("mh", "output") if kind == "houlsby" else ("output",)
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=seed, adapters={task: TINY.hidden_size // rf},
                             adapter_sites=sites)
    spec = vo.AdapterSpec(task, "swish" if kind == "houlsby" else "relu", sites)
    pooled, logits, loss, grads = _oracle_step(sd, TINY, task, batch, adapter=spec)
    assert np.allclose(pooled.detach().numpy(), g["pooled"], atol=2e-6)
    assert np.allclose(logits.detach().numpy(), g["logits"], atol=5e-6)
    trainable = set(g["trainable"].tolist())
    # train_adapter freezes ViltModel only: the adapter + ALL task heads stay trainable
    assert all((".adapters." in n) or n.startswith("task_layer.") for n in trainable)
    worst = compare_grads(g, grads, rtol_norm=1e-4, tol_elem=2e-4)
    print("worst grad", worst)


def test_tiny_ewc_matches_reference():
    g = load("tiny_ewc_snli-ve")
    seed = int(g["seed"])
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=seed)
    sizes = g["batch_sizes"].tolist()
    batch_grads = []
    for i, b in enumerate(sizes):
        batch = vo.synth_batch("snli-ve", b, TINY, T=TINY_T, image_hw=TINY_HW, seed=seed + i, masked=True)
        assert np.array_equal(g[f"b{i}_input_ids"], batch["input_ids"].numpy())
        _, _, _, grads = _oracle_step(sd, TINY, "snli-ve", batch)
        batch_grads.append({k[len("vilt_encoder."):]: v for k, v in grads.items()
                            if k.startswith("vilt_encoder.") and v is not None})
    fisher = vo.fisher_from_batch_grads(batch_grads, sizes)
    n_checked = 0
    fscale = max(float(g[k]) for k in g.files if k.startswith("fisher_norm/"))
    for key in g.files:
        if not key.startswith("fisher_norm/"):
            continue
        name = key[len("fisher_norm/"):]
        ref_norm = float(g[key])
        if ref_norm < 1e-10 * fscale:       # squared rounding noise of an analytically-zero gradient
            assert fisher[name].norm().item() < 1e-9 * fscale, name
            n_checked += 1
            continue
        assert abs(fisher[name].norm().item() - ref_norm) <= 2e-4 * max(ref_norm, 1e-20), name
        if "fisher/" + name in g.files:
            ref = torch.from_numpy(g["fisher/" + name])
            assert (fisher[name] - ref).abs().max().item() <= 5e-4 * max(ref.abs().max().item(), 1e-20), name
        n_checked += 1
    assert n_checked > 30
    # penalty and its gradient at the perturbed parameters
    gen = torch.Generator().manual_seed(seed + 99)
    theta_star = {k[len("vilt_encoder."):]: v for k, v in sd.items() if k.startswith("vilt_encoder.")}
    theta = {}
    for n, p in theta_star.items():         # same order as named_parameters() of the reference encoder
        theta[n] = (p + 0.01 * torch.randn(p.shape, generator=gen)).requires_grad_(True)
    loss = vo.ewc_penalty(theta, theta_star, fisher, float(g["ewc_loss_weight"]))
    loss.backward()
    assert abs(loss.item() - float(g["ewc_loss"])) <= 5e-4 * float(g["ewc_loss"])
    compare_grads(g, {n: t.grad for n, t in theta.items()}, rtol_norm=5e-4, tol_elem=1e-3)


@pytest.mark.parametrize("task,seed,masked", [("vqa", 42, False), ("nlvr2", 43, True)])
def test_base_config_matches_reference(task, seed, masked):
    """ViLT-base geometry (BASELINE.json configs[0] shape at B=2): 40 tokens + 14x14 patches."""
    g = load(f"base_{task}")
    batch = regen_batch(g, task, BASE, 40, BASE_HW, 2, seed, masked)
    sd = vo.synth_state_dict(BASE, ALL_TASKS, seed=seed)
    torch.set_num_threads(8)
    pooled, logits, loss, grads = _oracle_step(sd, BASE, task, batch)
    assert np.allclose(pooled.detach().numpy(), g["pooled"], atol=5e-6)
    assert np.allclose(logits.detach().numpy(), g["logits"], atol=2e-5)
    assert abs(loss.item() - float(g["loss"])) <= 5e-6 * abs(float(g["loss"]))
    worst = compare_grads(g, grads, rtol_norm=2e-4, tol_elem=5e-4)
    print("worst grad", worst)


TINY_BERT = vo.BertDims(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=200,
                        max_position_embeddings=16)


@pytest.mark.parametrize("task,seed", [("vcr", 400), ("nlvr2", 401), ("vqa", 402)])
def test_tiny_viltbert_matches_reference(task, seed):
    """ViLT-BERT (src/modeling/viltbert.py): frozen BERT features -> ViltModel(inputs_embeds=...)."""
    g = load(f"tiny_viltbert_{task}")
    batch = regen_batch(g, task, TINY, TINY_T, TINY_HW, 3, seed, True)
    sd = vo.synth_viltbert_state_dict(TINY, TINY_BERT, ALL_TASKS, seed=seed, **fixture_scales(g))
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pooled, logits = vo.viltbert_learner_forward(params, TINY, TINY_BERT, task, batch)
    loss = vo.task_loss(task, logits, batch["target"])
    loss.backward()
    assert np.allclose(pooled.detach().numpy(), g["pooled"], atol=2e-6)
    assert np.allclose(logits.detach().numpy(), g["logits"], atol=5e-6)
    assert abs(loss.item() - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    ids = batch["input_ids"] if batch["input_ids"].dim() == 2 else batch["input_ids"][:, 0]
    am = batch["attention_mask"] if batch["attention_mask"].dim() == 2 else batch["attention_mask"][:, 0]
    tt = batch["token_type_ids"] if batch["token_type_ids"].dim() == 2 else batch["token_type_ids"][:, 0]
    with torch.no_grad():
        feats = vo.bert_forward(sd, TINY_BERT, ids, am, tt)
    assert np.allclose(feats.numpy(), g["bert_hidden"], atol=5e-6)
    grads = {k: v.grad for k, v in params.items()}
    # what the reference leaves without a gradient: all of BERT (no_grad), ViLT's word table (inputs_embeds path),
    # the heads of the other tasks
    no_grad = set(g["no_grad"].tolist())
    assert all(n.startswith("viltbert_encoder.bert.") or n.endswith("text_embeddings.word_embeddings.weight")
               or n.startswith("task_layer.") for n in no_grad)
    for n in no_grad:
        assert grads[n] is None, n
    # VCR's near-uniform logits give gradients ~1e-5 of the weights' scale: fp32 summation order shows at 3e-4
    worst = compare_grads(g, grads, rtol_norm=2e-4, tol_elem=5e-4)
    print("worst grad", worst)


def test_bert_base_matches_reference():
    """bert-base geometry, eval mode: last_hidden_state of the vendored BertModel (random init)."""
    from tests.golden_util import grad_sample_index
    g = load("base_bert_hidden")
    bd = vo.BertDims()
    sd = vo.synth_bert_state_dict(bd, seed=int(g["seed"]), prefix="")
    ids, am, tt = (torch.from_numpy(g[k]) for k in ("in_input_ids", "in_attention_mask", "in_token_type_ids"))
    with torch.no_grad():
        h = vo.bert_forward(sd, bd, ids, am, tt, prefix="")
    assert abs(h.norm().item() - float(g["hidden_norm"])) <= 1e-5 * float(g["hidden_norm"])
    assert np.allclose(h[:, 0].numpy(), g["hidden_cls"], atol=2e-5)
    assert np.allclose(h.flatten().numpy()[grad_sample_index(h.numel())], g["hidden_sample"], atol=2e-5)


RAGGED = [("tiny_ragged_snli-ve", "snli-ve", 4, 500), ("tiny_ragged_nlvr2", "nlvr2", 3, 501), ("tiny_ragged_vcr", "vcr", 3, 502)]


@pytest.mark.parametrize("tag,task,B,seed", RAGGED)
def test_tiny_padded_images_match_reference(tag, task, B, seed):
    """Images of different sizes padded to a common H x W with pixel_mask zeros: the variable-resolution
    visual_embed path (modeling_vilt.py:121-205) -- per-image position grids, masked padding rows."""
    g = load(tag)
    batch = regen_batch(g, task, TINY, TINY_T, (64, 80), B, seed, True)
    assert "pixel_mask" in batch
    assert np.array_equal(g["in_pixel_values"], batch["pixel_values"].numpy())
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=seed, **fixture_scales(g))
    pooled, logits, loss, grads = _oracle_step(sd, TINY, task, batch)
    assert np.allclose(pooled.detach().numpy(), g["pooled"], atol=2e-6)
    assert np.allclose(logits.detach().numpy(), g["logits"], atol=5e-6)
    assert abs(loss.item() - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    worst = compare_grads(g, grads, rtol_norm=2e-4, tol_elem=5e-4)
    print("worst grad", worst)


MAXLEN = [("tiny_maxlen_snli-ve", "snli-ve", 4, 600, (64, 80)), ("tiny_maxlen_nlvr2", "nlvr2", 3, 601, (64, 80)),
          ("tiny_maxlen_vcr", "vcr", 3, 602, (64, 80)), ("tiny_maxlen_full_vqa", "vqa", 3, 603, TINY_HW)]


@pytest.mark.parametrize("tag,task,B,seed,hw", MAXLEN)
def test_tiny_max_image_length_matches_reference(tag, task, B, seed, hw):
    """config.max_image_length > 0 (modeling_vilt.py:163-189): images with more valid patches than the cap keep a random
    subset drawn with torch.multinomial. The restatement makes the same draws in the same order, so under the fixture's
    seed it keeps the patches the unmodified reference kept -- single image, image pair (two passes) and four choices."""
    import dataclasses
    g = load(tag)
    batch = regen_batch(g, task, TINY, TINY_T, hw, B, seed, True)
    dims = dataclasses.replace(TINY, max_image_length=int(g["max_image_length"]))
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=seed, **fixture_scales(g))
    torch.manual_seed(seed)
    pooled, logits, loss, grads = _oracle_step(sd, dims, task, batch)
    assert np.allclose(pooled.detach().numpy(), g["pooled"], atol=2e-6)
    assert np.allclose(logits.detach().numpy(), g["logits"], atol=5e-6)
    assert abs(loss.item() - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    compare_grads(g, grads, rtol_norm=2e-4, tol_elem=5e-4)
    # a different seed keeps different patches: the fixture really depends on the draw
    torch.manual_seed(seed + 1)
    pooled2, _, _, _ = _oracle_step(sd, dims, task, batch)
    assert not np.allclose(pooled2.detach().numpy(), g["pooled"], atol=2e-6)


def test_base_padded_images_match_reference():
    """ViLT-base geometry, 384 x 640 padded batch (12 x 20 patch grid) with three image sizes."""
    g = load("base_ragged_vqa")
    batch = regen_batch(g, "vqa", BASE, 40, (384, 640), 3, 44, True)
    sd = vo.synth_state_dict(BASE, ALL_TASKS, seed=44)
    torch.set_num_threads(8)
    pooled, logits, loss, grads = _oracle_step(sd, BASE, "vqa", batch)
    assert np.allclose(pooled.detach().numpy(), g["pooled"], atol=5e-6)
    assert np.allclose(logits.detach().numpy(), g["logits"], atol=2e-5)
    assert abs(loss.item() - float(g["loss"])) <= 5e-6 * abs(float(g["loss"]))
    compare_grads(g, grads, rtol_norm=2e-4, tol_elem=5e-4)
