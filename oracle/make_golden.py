"""Generate tests/golden/*.npz by running the UNMODIFIED reference (CLiMB @ /root/reference, with
its vendored adapter-transformers fork) on CPU through the shims of oracle/ref_shim.py.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden            # writes tests/golden/

TEST INFRASTRUCTURE: only runs in the build container (the GPU box has no /root/reference). The
fixtures it writes are committed; tests/test_oracle_golden.py replays them against oracle/ (CPU)
and tests/test_gpu_parity.py against the CUDA path.

What is exercised in the reference (no restatement on this side of the comparison):
  * ViltContinualLearner.forward -> forward_single_image / forward_multi_images /
    forward_multi_choice (src/modeling/vilt.py:218-350) with process_inputs replaced by a function
    returning our pre-made tensors (the PIL / tokenizer path needs the bert vocab, not offline);
  * ViltModel (modeling_vilt.py) incl. visual_embed with its multinomial patch permutation;
  * the trainers' loss expressions (train_vqa.py:95,157; train_nlvr2.py:80,133) and autograd;
  * adapter-transformers' add_adapter / train_adapter / Stack forward for Houlsby and Pfeiffer;
  * EWC.compute_ewc_loss and EWC.save_task_parameters (src/cl_algorithms/ewc.py:28-87);
  * ViltBertContinualLearner / ViltBertEncoderWrapper (src/modeling/viltbert.py) with the vendored
    BertModel (dropouts set to 0), and BertModel alone at bert-base geometry.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.vilt_oracle import (TASK_SPECS, BertDims, ViltDims, bert_param_shapes, pad_batch_images, synth_batch,  # noqa: E402
                                synth_bert_state_dict,
                                synth_state_dict, synth_viltbert_state_dict)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

TINY = ViltDims(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                image_size=32, patch_size=16, vocab_size=200, max_position_embeddings=8)
TINY_HW = (48, 64)          # 3 x 4 patches: non-square, so the 2x2 position grid is interpolated
TINY_T = 8
BASE = ViltDims()           # ViltConfig defaults = dandelin/vilt-b32 geometry
BASE_HW = (448, 448)        # 14 x 14 patches -> L = 40 + 197 = 237 (BASELINE.json configs)
ALL_TASKS = ["vqa", "nlvr2", "snli-ve", "vcr"]


def build_reference_learner(dims: ViltDims, tasks, sd):
    from transformers import ViltConfig, ViltModel
    from modeling.vilt import ViltContinualLearner, ViltEncoderWrapper
    from configs.task_configs import task_configs
    cfg = ViltConfig(hidden_size=dims.hidden_size, num_hidden_layers=dims.num_hidden_layers,
                     num_attention_heads=dims.num_attention_heads, intermediate_size=dims.intermediate_size,
                     image_size=dims.image_size, patch_size=dims.patch_size, vocab_size=dims.vocab_size,
                     max_position_embeddings=dims.max_position_embeddings, max_image_length=dims.max_image_length)
    enc = ViltEncoderWrapper(ref_shim.StubProcessor(), ViltModel(cfg), torch.device("cpu"))
    learner = ViltContinualLearner(list(tasks), enc, dims.hidden_size, task_configs)
    missing, unexpected = learner.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing
    return learner


def reference_inputs(task, batch):
    """The dict ViltProcessor would have returned for this batch (src/modeling/vilt.py:83-96), laid
    out the way forward_multi_images / forward_multi_choice expect to reshape it (:281-289, :326-332)."""
    spec = TASK_SPECS[task]
    px = batch["pixel_values"]
    ids, am, tt = batch["input_ids"], batch["attention_mask"], batch["token_type_ids"]
    if spec["num_images"] > 1:
        px = px.flatten(0, 1)
    if spec["model_type"] == "multi-choice":
        ids, am, tt = ids.flatten(0, 1), am.flatten(0, 1), tt.flatten(0, 1)
    pm = batch.get("pixel_mask")
    pm = torch.ones(px.shape[0], px.shape[-2], px.shape[-1], dtype=torch.long) if pm is None else pm.reshape(-1, *pm.shape[-2:])
    return {"input_ids": ids, "attention_mask": am, "token_type_ids": tt, "pixel_values": px, "pixel_mask": pm}


def reference_step(learner, task, batch):
    """forward + trainer loss + backward, as TaskTrainer.train_step with optimizer=None does
    (train_vqa.py:135-174)."""
    enc_inputs = reference_inputs(task, batch)
    learner.vilt_encoder.process_inputs = lambda images, texts: enc_inputs
    n = batch["pixel_values"].shape[0]
    spec = TASK_SPECS[task]
    images = [[None] * spec["num_images"]] * n if spec["num_images"] > 1 else [None] * n
    texts = [[None] * spec["num_choices"]] * n if spec["model_type"] == "multi-choice" else [None] * n
    pooled, logits = learner(task_key=task, images=images, texts=texts)
    if task == "vqa":
        loss = torch.nn.BCEWithLogitsLoss(reduction="mean")(logits, batch["target"]) * batch["target"].shape[1]
    else:
        loss = torch.nn.CrossEntropyLoss()(logits, batch["target"])
    loss.backward()
    return pooled, logits, loss


def batch_arrays(batch, store_pixels=True):
    """Inputs come from synth_batch(seed) and are regenerated by the tests; the copy stored here
    guards that regeneration (pixels of the base config only as a checksum: they are 2.4 MB/image)."""
    out = {"in_" + k: v.numpy() for k, v in batch.items() if store_pixels or k != "pixel_values"}
    px = batch["pixel_values"].double()
    out["in_pixel_checksum"] = np.array([px.sum().item(), px.abs().sum().item(), (px * px).sum().item()])
    return out


GRAD_FULL_MAX = 20_000       # larger gradients are stored as a norm + a strided sample
GRAD_SAMPLES = 2048


def grad_sample_index(numel: int) -> np.ndarray:
    """Deterministic sample positions shared with the tests."""
    if numel <= GRAD_SAMPLES:
        return np.arange(numel)
    return (np.arange(GRAD_SAMPLES, dtype=np.int64) * (numel - 1)) // (GRAD_SAMPLES - 1)


def store_grad(out, name, g, full):
    out["gnorm/" + name] = np.float32(g.norm().item())
    if full and g.numel() <= GRAD_FULL_MAX:
        out["grad/" + name] = g.numpy()
    else:
        out["gsample/" + name] = g.flatten().numpy()[grad_sample_index(g.numel())].copy()


# VCR fixtures use scaled encoder-layer / head matrices: at the plain 0.02 init the four choices of a sample pool to the same
# vector and the fixture pins nothing about the multi-choice backward (see synth_state_dict)
VCR_SCALES = dict(layer_scale=6.0, head_scale=20.0)


def scales_for(task):
    return VCR_SCALES if task == "vcr" else dict(layer_scale=1.0, head_scale=1.0)


def run_task(dims, hw, T, task, B, seed, tag, masked, full_grads, image_sizes=None, max_image_length=-1):
    tasks = ALL_TASKS
    if max_image_length > 0:        # ViltConfig.max_image_length: random patch dropping (modeling_vilt.py:163-189)
        import dataclasses
        dims = dataclasses.replace(dims, max_image_length=max_image_length)
    sc = scales_for(task)
    sd = synth_state_dict(dims, tasks, seed=seed, **sc)
    learner = build_reference_learner(dims, tasks, sd)
    learner.train()
    # the only active dropout on the path is the VCR head's Dropout(0.1) (src/modeling/vilt.py:199-202;
    # ViltConfig dropouts are 0.0): switched off so that the fixture is deterministic
    learner.task_layer["vcr"][0].eval()
    batch = synth_batch(task, B, dims, T=T, image_hw=hw, seed=seed, masked=masked)
    if image_sizes is not None:       # images of different sizes padded to hw with pixel_mask zeros (visual_embed :149-193)
        batch = pad_batch_images(batch, image_sizes)
    torch.manual_seed(seed)         # visual_embed's multinomial permutation
    pooled, logits, loss = reference_step(learner, task, batch)
    out = dict(batch_arrays(batch, store_pixels=full_grads), pooled=pooled.detach().numpy(),
               logits=logits.detach().numpy(),
               loss=np.float32(loss.item()), seed=np.int64(seed), task=task, B=np.int64(B), T=np.int64(T),
               hw=np.array(hw), masked=np.int64(masked), layer_scale=np.float32(sc["layer_scale"]),
               head_scale=np.float32(sc["head_scale"]))
    if image_sizes is not None:
        out["image_sizes"] = np.array(image_sizes)
    if max_image_length > 0:
        out["max_image_length"] = np.int64(max_image_length)
    for n, p in learner.named_parameters():
        g = p.grad
        if g is None:
            continue
        store_grad(out, n, g, full_grads)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{tag}.npz"), **out)
    print(f"{tag}: loss={loss.item():.6f} pooled[0,:3]={pooled[0].flatten()[:3].tolist()}")


def run_adapter(kind, task, rf, seed, tag):
    from transformers.adapters import AdapterConfig
    dims = TINY
    r = dims.hidden_size // rf
    sites = ("mh", "output") if kind == "houlsby" else ("output",)
    sd = synth_state_dict(dims, ALL_TASKS, seed=seed, adapters={task: r}, adapter_sites=sites)
    base_sd = {k: v for k, v in sd.items() if ".adapters." not in k}
    learner = build_reference_learner(dims, ALL_TASKS, base_sd)
    cfg = AdapterConfig.load(kind).to_dict()            # AdapterHandler.__init__, adapters.py:38-50
    cfg["reduction_factor"] = rf
    learner.add_adapter(task, config=AdapterConfig.from_dict(cfg))
    missing, unexpected = learner.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    learner.train_adapter(task)                         # activate_adapter_for_training, adapters.py:58-61
    learner.set_active_adapters(task)
    learner.train()
    batch = synth_batch(task, 3, dims, T=TINY_T, image_hw=TINY_HW, seed=seed, masked=True)
    torch.manual_seed(seed)
    pooled, logits, loss = reference_step(learner, task, batch)
    out = dict(batch_arrays(batch), pooled=pooled.detach().numpy(), logits=logits.detach().numpy(),
               loss=np.float32(loss.item()), seed=np.int64(seed), task=task, kind=kind, rf=np.int64(rf))
    trainable = []
    for n, p in learner.named_parameters():
        if p.requires_grad:
            trainable.append(n)
        if p.grad is not None:
            store_grad(out, n, p.grad, True)
    out["trainable"] = np.array(trainable)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{tag}.npz"), **out)
    print(f"{tag}: loss={loss.item():.6f} trainable={len(trainable)}")


def run_ewc(seed, tag):
    """EWC.save_task_parameters (Fisher with the cumulative-gradient quirk) over three synthetic
    batches, then EWC.compute_ewc_loss + its gradient at perturbed parameters."""
    from cl_algorithms.ewc import EWC
    dims = TINY
    sd = synth_state_dict(dims, ALL_TASKS, seed=seed)
    learner = build_reference_learner(dims, ALL_TASKS, sd)
    learner.train()
    batches = [synth_batch("snli-ve", b, dims, T=TINY_T, image_hw=TINY_HW, seed=seed + i, masked=True)
               for i, b in enumerate([2, 3, 2])]
    for b in batches:
        b["raw_texts"] = [None] * b["pixel_values"].shape[0]

    class FakeTrainer:          # what EWC touches on a TaskTrainer (ewc.py:48-60)
        hparams = {"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8}
        batch2inputs_converter = None
        loss_criterion = None
        device = torch.device("cpu")

        def get_train_dataloader(self):
            class DL(list):
                dataset = [None] * sum(len(b["raw_texts"]) for b in batches)
            return DL(batches)

        def train_step(self, model, batch, optimizer=None, scheduler=None, ewc=None):
            _, logits, loss = reference_step(model, "snli-ve", batch)
            return loss, logits, None, None

    args = types.SimpleNamespace(ewc_fisher_sample_percentage=1.0, ewc_loss_weight=100.0)
    ewc = EWC(args)
    torch.manual_seed(seed)
    import cl_algorithms.ewc as ewc_mod
    ewc_mod.tqdm = lambda it, **k: it
    ewc.save_task_parameters("snli-ve", learner, FakeTrainer(), torch.device("cpu"))
    out = {"seed": np.int64(seed), "ewc_loss_weight": np.float32(100.0), "batch_sizes": np.array([2, 3, 2])}
    for i, b in enumerate(batches):
        for k, v in b.items():
            if k != "raw_texts":
                out[f"b{i}_{k}"] = v.numpy()
    for n, f in ewc.fisher_dict["snli-ve"].items():
        out["fisher_norm/" + n] = np.float32(f.norm().item())
        if f.numel() <= GRAD_FULL_MAX:
            out["fisher/" + n] = f.numpy()
        else:
            out["fisher_sample/" + n] = f.flatten().numpy()[grad_sample_index(f.numel())].copy()
    # perturb parameters, then penalty + gradient
    g = torch.Generator().manual_seed(seed + 99)
    with torch.no_grad():
        for n, p in learner.get_encoder().named_parameters():
            p.add_(0.01 * torch.randn(p.shape, generator=g))     # tests regenerate the same deltas
    learner.zero_grad()
    random.seed(0)
    key, loss = ewc.compute_ewc_loss(learner)
    loss.backward()
    out["ewc_loss"] = np.float32(loss.item())
    for n, p in learner.get_encoder().named_parameters():
        if p.grad is not None:
            store_grad(out, n, p.grad, True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{tag}.npz"), **out)
    print(f"{tag}: task={key} ewc_loss={loss.item():.6f}")


TINY_BERT = BertDims(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=200,
                     max_position_embeddings=16)


def build_reference_viltbert(dims: ViltDims, bdims: BertDims, tasks, sd):
    """ViltBertContinualLearner over ViltBertEncoderWrapper(processor, ViltModel, BertModel) (viltbert.py:31-57,
    171-201). BertConfig dropouts are set to 0 so that the fixture is deterministic in train mode (the
    reference leaves BERT's 0.1 dropouts live inside no_grad: a documented quirk, not a parity target)."""
    from transformers import BertConfig, BertModel, ViltConfig, ViltModel
    from modeling.viltbert import ViltBertContinualLearner, ViltBertEncoderWrapper
    from configs.task_configs import task_configs
    cfg = ViltConfig(hidden_size=dims.hidden_size, num_hidden_layers=dims.num_hidden_layers,
                     num_attention_heads=dims.num_attention_heads, intermediate_size=dims.intermediate_size,
                     image_size=dims.image_size, patch_size=dims.patch_size, vocab_size=dims.vocab_size,
                     max_position_embeddings=dims.max_position_embeddings)
    bcfg = BertConfig(hidden_size=bdims.hidden_size, num_hidden_layers=bdims.num_hidden_layers,
                      num_attention_heads=bdims.num_attention_heads, intermediate_size=bdims.intermediate_size,
                      vocab_size=bdims.vocab_size, max_position_embeddings=bdims.max_position_embeddings,
                      type_vocab_size=bdims.type_vocab_size, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    enc = ViltBertEncoderWrapper(ref_shim.StubProcessor(), ViltModel(cfg), BertModel(bcfg), torch.device("cpu"))
    learner = ViltBertContinualLearner(list(tasks), enc, dims.hidden_size, task_configs)
    missing, unexpected = learner.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing
    return learner


def run_viltbert(task, B, seed, tag, masked=True):
    """ViLT-BERT (BASELINE.json config 5's encoder) on the tiny geometry: forward_single_image /
    forward_multi_images / forward_multi_choice of viltbert.py:231-345, loss, backward."""
    dims, bdims = TINY, TINY_BERT
    sc = scales_for(task)
    sd = synth_viltbert_state_dict(dims, bdims, ALL_TASKS, seed=seed, **sc)
    learner = build_reference_viltbert(dims, bdims, ALL_TASKS, sd)
    learner.train()
    learner.task_layer["vcr"][0].eval()
    batch = synth_batch(task, B, dims, T=TINY_T, image_hw=TINY_HW, seed=seed, masked=masked)
    enc_inputs = reference_inputs(task, batch)
    learner.viltbert_encoder.process_inputs = lambda images, texts: dict(enc_inputs)
    n = batch["pixel_values"].shape[0]
    spec = TASK_SPECS[task]
    images = [[None] * spec["num_images"]] * n if spec["num_images"] > 1 else [None] * n
    texts = [[None] * spec["num_choices"]] * n if spec["model_type"] == "multi-choice" else [None] * n
    torch.manual_seed(seed)
    pooled, logits = learner(task_key=task, images=images, texts=texts)
    if task == "vqa":
        loss = torch.nn.BCEWithLogitsLoss(reduction="mean")(logits, batch["target"]) * batch["target"].shape[1]
    else:
        loss = torch.nn.CrossEntropyLoss()(logits, batch["target"])
    loss.backward()
    with torch.no_grad():           # the frozen BERT's features for the first text of every sample
        ids = batch["input_ids"] if batch["input_ids"].dim() == 2 else batch["input_ids"][:, 0]
        am = batch["attention_mask"] if batch["attention_mask"].dim() == 2 else batch["attention_mask"][:, 0]
        tt = batch["token_type_ids"] if batch["token_type_ids"].dim() == 2 else batch["token_type_ids"][:, 0]
        feats = learner.viltbert_encoder.get_bert_outputs(input_ids=ids, attention_mask=am, token_type_ids=tt)
    out = dict(batch_arrays(batch), pooled=pooled.detach().numpy(), logits=logits.detach().numpy(), bert_hidden=feats.numpy(),
               loss=np.float32(loss.item()), seed=np.int64(seed), task=task, B=np.int64(B), T=np.int64(TINY_T),
               hw=np.array(TINY_HW), masked=np.int64(masked), layer_scale=np.float32(sc["layer_scale"]),
               head_scale=np.float32(sc["head_scale"]))
    no_grad = []
    for n_, p in learner.named_parameters():
        if p.grad is None:
            no_grad.append(n_)
        else:
            store_grad(out, n_, p.grad, True)
    out["no_grad"] = np.array(no_grad)
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{tag}.npz"), **out)
    print(f"{tag}: loss={loss.item():.6f} params without grad={len(no_grad)}")


def run_bert_base(seed, tag):
    """bert-base-uncased geometry (random init, eval mode): last_hidden_state of the unmodified BertModel on
    a masked (B=2, T=40) batch, stored as a strided sample + norm."""
    from transformers import BertConfig, BertModel
    bdims = BertDims()
    sd = synth_bert_state_dict(bdims, seed=seed, prefix="")
    model = BertModel(BertConfig())
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
    model.eval()
    batch = synth_batch("snli-ve", 2, ViltDims(), T=40, image_hw=(32, 32), seed=seed, masked=True)
    with torch.no_grad():
        h = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                  token_type_ids=batch["token_type_ids"]).last_hidden_state
    out = {"in_input_ids": batch["input_ids"].numpy(), "in_attention_mask": batch["attention_mask"].numpy(),
           "in_token_type_ids": batch["token_type_ids"].numpy(), "seed": np.int64(seed),
           "hidden_norm": np.float32(h.norm().item()), "hidden_cls": h[:, 0].numpy(),
           "hidden_sample": h.flatten().numpy()[grad_sample_index(h.numel())].copy()}
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{tag}.npz"), **out)
    print(f"{tag}: |h|={h.norm().item():.4f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-base", action="store_true")
    ap.add_argument("--only-viltbert", action="store_true", help="regenerate the ViLT-BERT fixtures only")
    ap.add_argument("--only-ragged", action="store_true", help="regenerate the padded-image (pixel_mask) fixtures only")
    ap.add_argument("--only-vcr", action="store_true", help="regenerate the three VCR (multi-choice) fixtures only")
    ap.add_argument("--only-maxlen", action="store_true", help="regenerate the max_image_length (patch dropping) fixtures only")
    a = ap.parse_args()
    ref_shim.install()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    # config.max_image_length > 0: 4 x 5 patch grid capped at 9 rows -- larger images keep a random subset (the seed set in
    # run_task fixes it), smaller ones are padded; single image, image pair (two encoder passes = two rounds of draws) and
    # four choices (four passes over the same pixels, each with its own subset); one batch without any pixel_mask padding
    run_task(TINY, (64, 80), TINY_T, "snli-ve", B=4, seed=600, tag="tiny_maxlen_snli-ve", masked=True, full_grads=True,
             image_sizes=[(64, 80), (48, 64), (32, 64), (64, 32)], max_image_length=9)
    run_task(TINY, (64, 80), TINY_T, "nlvr2", B=3, seed=601, tag="tiny_maxlen_nlvr2", masked=True, full_grads=True,
             image_sizes=[(64, 80), (48, 48), (32, 64), (64, 48), (16, 80), (48, 80)], max_image_length=9)
    run_task(TINY, (64, 80), TINY_T, "vcr", B=3, seed=602, tag="tiny_maxlen_vcr", masked=True, full_grads=True,
             image_sizes=[(48, 80), (64, 64), (32, 48)], max_image_length=9)
    run_task(TINY, TINY_HW, TINY_T, "vqa", B=3, seed=603, tag="tiny_maxlen_full_vqa", masked=True, full_grads=True, max_image_length=7)
    if a.only_maxlen:
        return
    if a.only_vcr:
        run_task(TINY, (64, 80), TINY_T, "vcr", B=3, seed=502, tag="tiny_ragged_vcr", masked=True, full_grads=True,
                 image_sizes=[(48, 80), (64, 64), (32, 48)])
        run_viltbert("vcr", 3, 400, "tiny_viltbert_vcr")
        run_task(TINY, TINY_HW, TINY_T, "vcr", B=3, seed=103, tag="tiny_vcr", masked=True, full_grads=True)
        return
    # padded batches: images of different sizes (pixel_mask zeros), tiny geometry (patch 16, padded to 64 x 80 = 4 x 5
    # patches) for single-image, image-pair and four-choice tasks, and the ViLT-base geometry (patch 32, 384 x 640)
    run_task(TINY, (64, 80), TINY_T, "snli-ve", B=4, seed=500, tag="tiny_ragged_snli-ve", masked=True, full_grads=True,
             image_sizes=[(64, 80), (48, 64), (32, 80), (64, 32)])
    run_task(TINY, (64, 80), TINY_T, "nlvr2", B=3, seed=501, tag="tiny_ragged_nlvr2", masked=True, full_grads=True,
             image_sizes=[(64, 80), (48, 48), (32, 64), (64, 48), (16, 80), (48, 80)])
    run_task(TINY, (64, 80), TINY_T, "vcr", B=3, seed=502, tag="tiny_ragged_vcr", masked=True, full_grads=True,
             image_sizes=[(48, 80), (64, 64), (32, 48)])
    if not a.skip_base:
        run_task(BASE, (384, 640), 40, "vqa", B=3, seed=44, tag="base_ragged_vqa", masked=True, full_grads=False,
                 image_sizes=[(384, 640), (384, 512), (352, 576)])
    if a.only_ragged:
        return
    run_viltbert("vcr", 3, 400, "tiny_viltbert_vcr")
    run_viltbert("nlvr2", 3, 401, "tiny_viltbert_nlvr2")
    run_viltbert("vqa", 3, 402, "tiny_viltbert_vqa")
    if not a.skip_base:
        run_bert_base(42, "base_bert_hidden")
    if a.only_viltbert:
        return
    for i, task in enumerate(ALL_TASKS):
        run_task(TINY, TINY_HW, TINY_T, task, B=3, seed=100 + i, tag=f"tiny_{task}", masked=True, full_grads=True)
    run_adapter("houlsby", "nlvr2", 4, 200, "tiny_adapter_houlsby_nlvr2")
    run_adapter("pfeiffer", "vqa", 2, 201, "tiny_adapter_pfeiffer_vqa")
    run_ewc(300, "tiny_ewc_snli-ve")
    if not a.skip_base:
        run_task(BASE, BASE_HW, 40, "vqa", B=2, seed=42, tag="base_vqa", masked=False, full_grads=False)
        run_task(BASE, BASE_HW, 40, "nlvr2", B=2, seed=43, tag="base_nlvr2", masked=True, full_grads=False)


if __name__ == "__main__":
    main()
