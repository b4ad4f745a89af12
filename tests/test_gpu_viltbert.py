"""GPU parity of the ViLT-BERT path (B200ViltBertContinualLearner: climb_bert_forward -> climb_vilt_forward /
backward with inputs_embeds) against the golden vectors of the UNMODIFIED reference
(src/modeling/viltbert.py + vendored BertModel, oracle/make_golden.py) and the CPU oracle.
Tolerances as in tests/test_gpu_parity.py (bf16 tensor-core operands): outputs 2e-2, gradients 6e-2."""
import math

import numpy as np
import pytest
import torch

from oracle import vilt_oracle as vo
from tests.golden_util import ALL_TASKS, TINY, TINY_HW, TINY_T, fixture_scales, grad_sample_index, load, regen_batch
from tests.test_gpu_parity import TOL_OUT, _check_grads, _encodings, _rel, check_outputs, gate

pytestmark = pytest.mark.gpu

TINY_BERT = vo.BertDims(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=200,
                        max_position_embeddings=16)


def _bert_cfg(b, p=0.0):
    from climb_b200.modeling import B200BertConfig
    return B200BertConfig(vocab_size=b.vocab_size, hidden_size=b.hidden_size, num_hidden_layers=b.num_hidden_layers,
                          num_attention_heads=b.num_attention_heads, intermediate_size=b.intermediate_size,
                          max_position_embeddings=b.max_position_embeddings, type_vocab_size=b.type_vocab_size,
                          hidden_dropout_prob=p, attention_probs_dropout_prob=p)


def _build(dims, bdims, tasks, sd, p=0.0):
    from climb_b200.modeling import (B200BertModel, B200ViltBertContinualLearner, B200ViltBertEncoderWrapper, B200ViltConfig,
                                     B200ViltModel)
    cfg = B200ViltConfig(hidden_size=dims.hidden_size, num_hidden_layers=dims.num_hidden_layers,
                         num_attention_heads=dims.num_attention_heads, intermediate_size=dims.intermediate_size,
                         image_size=dims.image_size, patch_size=dims.patch_size, vocab_size=dims.vocab_size,
                         max_position_embeddings=dims.max_position_embeddings)
    dev = torch.device("cuda")
    enc = B200ViltBertEncoderWrapper(None, B200ViltModel(cfg), B200BertModel(_bert_cfg(bdims, p)), dev)
    learner = B200ViltBertContinualLearner(list(tasks), enc, dims.hidden_size, vo.TASK_SPECS)
    missing, unexpected = learner.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in m for m in missing), missing
    return learner.to(dev)


@pytest.mark.parametrize("task,seed", [("vcr", 400), ("nlvr2", 401), ("vqa", 402)])
def test_tiny_viltbert_vs_reference_golden(task, seed):
    g = load(f"tiny_viltbert_{task}")
    batch = regen_batch(g, task, TINY, TINY_T, TINY_HW, 3, seed, True)
    sd = vo.synth_viltbert_state_dict(TINY, TINY_BERT, ALL_TASKS, seed=seed, **fixture_scales(g))
    learner = _build(TINY, TINY_BERT, ALL_TASKS, sd)
    assert set(learner.state_dict().keys()) >= set(sd.keys())          # checkpoint keys of the reference learner
    dev = torch.device("cuda")
    learner.train()
    learner.task_layer["vcr"][0].eval()
    enc = _encodings(task, batch, dev)
    # frozen BERT features (first text of every sample)
    ids = batch["input_ids"] if batch["input_ids"].dim() == 2 else batch["input_ids"][:, 0]
    am = batch["attention_mask"] if batch["attention_mask"].dim() == 2 else batch["attention_mask"][:, 0]
    tt = batch["token_type_ids"] if batch["token_type_ids"].dim() == 2 else batch["token_type_ids"][:, 0]
    feats = learner.get_encoder().get_bert_outputs(input_ids=ids.to(dev), attention_mask=am.to(dev), token_type_ids=tt.to(dev))
    e_f = _rel(feats, g["bert_hidden"])
    pooled, logits = learner.forward_tensors(task, enc)
    target = batch["target"].to(dev)
    loss = (torch.nn.BCEWithLogitsLoss()(logits, target) * target.shape[1]) if task == "vqa" else torch.nn.CrossEntropyLoss()(logits, target)
    loss.backward()
    gate(f"tiny_viltbert_{task}/bert", e_f, TOL_OUT)
    check_outputs(f"tiny_viltbert_{task}", pooled, logits, loss.item(), g["pooled"], g["logits"], g["loss"])
    # parameters the reference leaves without a gradient stay without one: all of BERT, ViLT's word table, other heads
    grads = {n: p.grad for n, p in learner.named_parameters()}
    for n in g["no_grad"].tolist():
        assert grads[n] is None, n
    _check_grads(g, learner, f"tiny_viltbert_{task}")


def test_bert_base_vs_reference_golden():
    """bert-base geometry (12 layers, d = 768, T = 40, masked) against the vendored BertModel's output."""
    from climb_b200.modeling import B200BertModel
    g = load("base_bert_hidden")
    bd = vo.BertDims()
    sd = vo.synth_bert_state_dict(bd, seed=int(g["seed"]), prefix="")
    dev = torch.device("cuda")
    bert = B200BertModel(_bert_cfg(bd, 0.1))
    missing, unexpected = bert.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing)
    bert.to(dev).eval()                      # eval: BertConfig's 0.1 dropouts are off, as in the golden run
    ids, am, tt = (torch.from_numpy(g[k]).to(dev) for k in ("in_input_ids", "in_attention_mask", "in_token_type_ids"))
    h = bert(input_ids=ids, attention_mask=am, token_type_ids=tt).last_hidden_state
    assert not h.requires_grad
    hc = h.float().cpu()
    e_cls = _rel(hc[:, 0], g["hidden_cls"])
    e_s = _rel(hc.flatten()[torch.from_numpy(grad_sample_index(hc.numel()))], g["hidden_sample"])
    print(f"bert-base: cls rel {e_cls:.3e} sample rel {e_s:.3e} norm {hc.norm().item():.3f} vs {float(g['hidden_norm']):.3f}")
    gate("base_bert/cls", e_cls, TOL_OUT)
    gate("base_bert/sample", e_s, TOL_OUT)
    assert abs(hc.norm().item() - float(g["hidden_norm"])) <= 5e-3 * float(g["hidden_norm"])


def test_nlvr2_bert_features_reused_across_images():
    """forward_multi_images runs BERT once per text; the result equals the per-image evaluation of the oracle."""
    sd = vo.synth_viltbert_state_dict(TINY, TINY_BERT, ALL_TASKS, seed=5)
    learner = _build(TINY, TINY_BERT, ALL_TASKS, sd)
    batch = vo.synth_batch("nlvr2", 4, TINY, T=TINY_T, image_hw=TINY_HW, seed=55, masked=True)
    learner.eval()
    with torch.no_grad():
        pooled, logits = learner.forward_tensors("nlvr2", _encodings("nlvr2", batch, torch.device("cuda")))
    ref_p, ref_l = vo.viltbert_learner_forward(sd, TINY, TINY_BERT, "nlvr2", batch)
    check_outputs("viltbert_fresh", pooled, logits, 0.0, ref_p, ref_l, None)


def test_attention_dropout_mask_statistics():
    """dropout(softmax) V with uniform probabilities and V = I exposes the mask itself: entries are 0 or
    1 / (L (1 - p)); the kept fraction is 1 - p; the mask is a pure function of the seed."""
    from climb_b200 import _lib as L
    B, Lq, H, p = 3, 64, 2, 0.25
    qkv = torch.zeros(B, Lq, 3 * H * 64, device="cuda", dtype=torch.bfloat16)
    eye = torch.eye(64, device="cuda", dtype=torch.bfloat16)
    for h in range(H):
        qkv[:, :, (2 * H + h) * 64:(2 * H + h + 1) * 64] = eye            # V_h = I (L = dh = 64)
    ctx1, lse = L.attention_fwd(qkv, None, B, Lq, H, 0.125, p_drop=p, seed=1234)
    ctx2, _ = L.attention_fwd(qkv, None, B, Lq, H, 0.125, p_drop=p, seed=1234)
    ctx3, _ = L.attention_fwd(qkv, None, B, Lq, H, 0.125, p_drop=p, seed=99)
    assert torch.equal(ctx1, ctx2) and not torch.equal(ctx1, ctx3)
    kept_val = 1.0 / (Lq * (1.0 - p))
    c = ctx1.float()
    is_zero, is_kept = c == 0, (c - kept_val).abs() <= 2 ** -8 * kept_val
    assert bool((is_zero | is_kept).all())
    frac = is_kept.float().mean().item()
    n = c.numel()
    assert abs(frac - (1 - p)) < 5 * math.sqrt(p * (1 - p) / n), frac       # 5 sigma
    assert torch.allclose(lse, torch.full_like(lse, math.log(Lq)), atol=1e-4)   # the normaliser ignores the mask
    # every (b, h) draws its own mask
    m = is_kept.view(B, Lq, H, 64)
    assert not torch.equal(m[0, :, 0], m[0, :, 1]) and not torch.equal(m[0, :, 0], m[1, :, 0])


def test_bert_train_mode_dropout_is_unbiased_and_seeded():
    """Train mode keeps BertConfig's dropouts live (reference quirk): features vary with the seed, are
    reproducible under torch.manual_seed, and average to the eval features."""
    from climb_b200 import _lib as L
    x = torch.randn(1 << 16, device="cuda")
    r = torch.randn(1 << 16, device="cuda")
    y = torch.empty_like(x)
    L.check(L.climb_dropout_add(L.ptr(x), L.ptr(r), L.ptr(y), x.numel(), 0.1, 7, L.stream()))
    d = y - r
    kept = d != 0
    assert abs(kept.float().mean().item() - 0.9) < 5 * math.sqrt(0.09 / x.numel())
    assert torch.allclose(d[kept], x[kept] / 0.9, rtol=1e-4, atol=1e-5)       # (x / 0.9 + r) - r: rounding of the add
    from climb_b200.modeling import B200BertModel
    bert = B200BertModel(_bert_cfg(TINY_BERT, 0.1)).to("cuda")
    ids = torch.randint(1, 200, (4, 8), device="cuda")
    am = torch.ones(4, 8, dtype=torch.long, device="cuda")
    bert.eval()
    h_eval = bert(input_ids=ids, attention_mask=am).last_hidden_state
    assert torch.equal(h_eval, bert(input_ids=ids, attention_mask=am).last_hidden_state)
    bert.train()
    torch.manual_seed(3)
    h1 = bert(input_ids=ids, attention_mask=am).last_hidden_state
    torch.manual_seed(3)
    h2 = bert(input_ids=ids, attention_mask=am).last_hidden_state
    h3 = bert(input_ids=ids, attention_mask=am).last_hidden_state
    assert torch.equal(h1, h2) and not torch.equal(h1, h3) and not torch.equal(h1, h_eval)
    assert _rel(h1, h_eval) < 0.9          # perturbed, not destroyed
