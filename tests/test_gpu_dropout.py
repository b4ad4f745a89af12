"""Dropout INSIDE the ViLT encoder (ViltConfig.hidden_dropout_prob / attention_probs_dropout_prob > 0, train mode): the five
sites of modeling_vilt.py -- :303 text embeddings, :201 image embeddings, :374 attention probabilities, :410 self-output
dense, :482 output dense -- forward AND backward, against the CPU oracle fed with the SAME masks.

The masks are counter-based (Philox, a pure function of the forward's seed and the element index: csrc/dropout.cu,
attention_tc.cu), so nothing is stored for the backward; climb_dropout_keep_mask / climb_attention_dropout_keep_mask
materialise them for the oracle. torch's own dropout uses a different generator, so parity with the reference is parity
of the ARITHMETIC under given masks plus the statistics of the masks themselves.
"""
import pytest
import torch

from oracle import vilt_oracle as vo
from tests.golden_util import ALL_TASKS, TINY, TINY_T
from tests.test_gpu_parity import TOL_GRAD, TOL_OUT, _encodings, _rel, gate

pytestmark = pytest.mark.gpu
P_HID, P_ATT = 0.1, 0.15


def _build(dims, sd, p_hid, p_att):
    from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel
    cfg = B200ViltConfig(hidden_size=dims.hidden_size, num_hidden_layers=dims.num_hidden_layers,
                         num_attention_heads=dims.num_attention_heads, intermediate_size=dims.intermediate_size,
                         image_size=dims.image_size, patch_size=dims.patch_size, vocab_size=dims.vocab_size,
                         max_position_embeddings=dims.max_position_embeddings, hidden_dropout_prob=p_hid,
                         attention_probs_dropout_prob=p_att)
    dev = torch.device("cuda")
    learner = B200ViltContinualLearner(list(ALL_TASKS), B200ViltEncoderWrapper(None, B200ViltModel(cfg), dev), dims.hidden_size,
                                       vo.TASK_SPECS)
    learner.load_state_dict(sd, strict=False)
    return learner.to(dev)


def _masks(seed, dims, B, L, p_hid, p_att):
    """The keep-factor tensors of every site, as the kernels generate them, on the CPU for the oracle."""
    from climb_b200 import _lib
    dev = torch.device("cuda")
    d, H = dims.hidden_size, dims.num_attention_heads
    out = {}

    def hidden(layer, site):
        t = torch.empty(B, L, d, dtype=torch.float32, device=dev)
        _lib.check(_lib.climb_dropout_keep_mask(_lib.ptr(t), t.numel(), p_hid, _lib.climb_dropout_site_seed(seed, layer, site),
                                                _lib.stream()))
        return t.cpu()

    if p_hid > 0:
        out["embed"] = hidden(-1, 1)
    for i in range(dims.num_hidden_layers):
        if p_att > 0:
            t = torch.empty(B, H, L, L, dtype=torch.float32, device=dev)
            _lib.check(_lib.climb_attention_dropout_keep_mask(_lib.ptr(t), B, H, L, p_att, _lib.climb_dropout_site_seed(seed, i, 0),
                                                             _lib.stream()))
            out[("attn", i)] = t.cpu()
        if p_hid > 0:
            out[("self_out", i)] = hidden(i, 1)
            out[("out", i)] = hidden(i, 2)
    return out


def _drawn_seed(torch_seed):
    """What B200ViltModel draws from torch's CPU generator for a forward issued right after torch.manual_seed(torch_seed)."""
    torch.manual_seed(torch_seed)
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _worst_grad(learner, params):
    named = dict(learner.named_parameters())
    gscale = max(v.grad.norm().item() for v in params.values() if v.grad is not None)
    return max(((named[n].grad.float().cpu() - p.grad).norm().item() / max(p.grad.norm().item(), 0.02 * gscale), n)
               for n, p in params.items() if p.grad is not None)


@pytest.mark.parametrize("image_hw,p_hid,p_att", [((48, 64), P_HID, P_ATT), ((192, 192), P_HID, P_ATT), ((48, 64), 0.0, 0.2),
                                                   ((48, 64), 0.2, 0.0)])
def test_encoder_dropout_forward_and_backward_vs_oracle_with_the_same_masks(image_hw, p_hid, p_att):
    """(192, 192) gives L = 8 + 1 + 144 = 153 tokens: two query / key tiles in the attention kernels."""
    dev = torch.device("cuda")
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=61)
    learner = _build(TINY, sd, p_hid, p_att)
    batch = vo.synth_batch("snli-ve", 3, TINY, T=TINY_T, image_hw=image_hw, seed=62, masked=True)
    enc = _encodings("snli-ve", batch, dev)
    enc.pop("pixel_mask")
    learner.train()
    seed = _drawn_seed(1234)
    torch.manual_seed(1234)
    pooled, logits = learner.forward_tensors("snli-ve", enc)
    loss = torch.nn.CrossEntropyLoss()(logits, batch["target"].to(dev))
    loss.backward()
    L = TINY_T + 1 + (image_hw[0] // TINY.patch_size) * (image_hw[1] // TINY.patch_size)
    masks = _masks(seed, TINY, 3, L, p_hid, p_att)
    for k, m in masks.items():                       # the masks themselves: values in {0, 1 / (1 - p)}, kept fraction ~ 1 - p
        p = p_att if (isinstance(k, tuple) and k[0] == "attn") else p_hid
        kept = (m != 0)
        assert torch.allclose(m[kept], torch.full_like(m[kept], 1.0 / (1.0 - p)))
        assert abs(kept.float().mean().item() - (1.0 - p)) < 0.02, (k, kept.float().mean().item())
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_p, ref_l = vo.learner_forward(params, TINY, "snli-ve", batch, dropout_masks=masks)
    ref_loss = vo.task_loss("snli-ve", ref_l, batch["target"])
    ref_loss.backward()
    key = f"dropout/L{L}_ph{p_hid}_pa{p_att}"
    gate(key + "/pooled", _rel(pooled, ref_p), TOL_OUT)
    gate(key + "/logits", _rel(logits, ref_l), TOL_OUT)
    worst = _worst_grad(learner, params)
    print("dropout worst gradient error", worst)
    gate(key + "/grad", worst[0], TOL_GRAD)
    # the dropout really happened: the no-dropout oracle is further away than the bf16 error (at random init the attention
    # is nearly uniform, so dropping probabilities alone moves the output least)
    clean_p, _ = vo.learner_forward(sd, TINY, "snli-ve", batch)
    assert _rel(pooled, clean_p) > (5 if p_hid > 0 else 2) * _rel(pooled, ref_p)


def test_eval_mode_and_zero_probability_switch_dropout_off_and_seeds_give_different_masks():
    dev = torch.device("cuda")
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=63)
    batch = vo.synth_batch("snli-ve", 2, TINY, T=TINY_T, image_hw=(48, 64), seed=64)
    enc = _encodings("snli-ve", batch, dev)
    with_drop = _build(TINY, sd, P_HID, P_ATT)
    without = _build(TINY, sd, 0.0, 0.0)
    with torch.no_grad():
        with_drop.eval()
        without.eval()
        p_eval, _ = with_drop.forward_tensors("snli-ve", enc)
        p_ref, _ = without.forward_tensors("snli-ve", enc)
        assert torch.equal(p_eval, p_ref)                       # eval: dropout is the identity, bit for bit
        with_drop.train()
        torch.manual_seed(1)
        a, _ = with_drop.forward_tensors("snli-ve", enc)
        torch.manual_seed(1)
        b, _ = with_drop.forward_tensors("snli-ve", enc)
        torch.manual_seed(2)
        c, _ = with_drop.forward_tensors("snli-ve", enc)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, p_ref)


def test_dropout_with_adapters_trains_the_adapter_through_the_masks():
    """Frozen base + Houlsby adapter with both dropouts on (mh adapter input = the dropped dense output, modeling_vilt.py:410-412)."""
    from climb_b200.modeling import AdapterSpec
    dev = torch.device("cuda")
    r = TINY.hidden_size // 4
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=65, adapters={"snli-ve": r})
    learner = _build(TINY, {k: v for k, v in sd.items() if ".adapters." not in k}, P_HID, P_ATT)
    spec = AdapterSpec.from_config("houlsby")
    spec.reduction_factor = 4
    learner.add_adapter("snli-ve", spec)
    learner.load_state_dict(sd, strict=False)
    learner.to(dev)
    learner.train_adapter("snli-ve")
    batch = vo.synth_batch("snli-ve", 3, TINY, T=TINY_T, image_hw=(48, 64), seed=66, masked=True)
    enc = _encodings("snli-ve", batch, dev)
    enc.pop("pixel_mask")
    seed = _drawn_seed(77)
    torch.manual_seed(77)
    learner.train()
    pooled, logits = learner.forward_tensors("snli-ve", enc)
    torch.nn.CrossEntropyLoss()(logits, batch["target"].to(dev)).backward()
    L = TINY_T + 1 + 12
    masks = _masks(seed, TINY, 3, L, P_HID, P_ATT)
    params = {k: v.clone().requires_grad_(".adapters." in k or k.startswith("task_layer.")) for k, v in sd.items()}
    ref_p, ref_l = vo.learner_forward(params, TINY, "snli-ve", batch, adapter=vo.AdapterSpec("snli-ve", "swish"), dropout_masks=masks)
    vo.task_loss("snli-ve", ref_l, batch["target"]).backward()
    gate("dropout/adapters/pooled", _rel(pooled, ref_p), TOL_OUT)
    gate("dropout/adapters/grad", _worst_grad(learner, params)[0], TOL_GRAD)
