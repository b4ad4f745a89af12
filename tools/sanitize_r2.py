"""Round-2 kernels and engine paths under compute-sanitizer (tiny shapes):
    compute-sanitizer --tool memcheck python tools/sanitize_r2.py
Covers: the fused adapter bottleneck kernel (forward / backward, r = 16 and 64, ragged row counts), the CTA-pair weight-gradient
kernel with the fused bias-gradient sums (colsum_a), ViLT-encoder dropout p > 0 (embeddings, attention probabilities, hidden)
forward + backward, VCR's shared image (image_repeat) on fixed and padded batches, max_image_length patch selection
(patch_select), the bf16x3 precision engine, an adapter step through the fused kernel inside the engine."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import _lib as L, ops  # noqa: E402
from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel  # noqa: E402

dev = torch.device("cuda")
P, S = L.ptr, L.stream()
g = torch.Generator().manual_seed(0)
bf = lambda *s: (torch.randn(*s, generator=g) * 0.1).to(dev).bfloat16()
f32 = lambda *s: torch.randn(*s, generator=g).to(dev)

# ---- fused adapter kernel ----
for M, d, r, act in ((300, 128, 16, L.EPI_SWISH), (129, 256, 64, L.EPI_RELU)):
    A, wd, wu, bd, bu = bf(M, d), bf(r, d), bf(d, r), f32(r), f32(d)
    pre, z = torch.empty(M, r, device=dev, dtype=torch.bfloat16), torch.empty(M, r, device=dev, dtype=torch.bfloat16)
    c, c2, cs = f32(M, d), torch.empty(M, d, device=dev, dtype=torch.bfloat16), torch.zeros(r, device=dev)
    L.check(L.climb_adapter_fused(0, M, d, r, act, P(A), P(wd), P(wu), P(bd), P(bu), P(pre), P(z), P(c), P(c), P(c2), None, S))
    L.check(L.climb_adapter_fused(1, M, d, r, act, P(A), P(wd), P(wu), None, None, P(pre), P(z), P(c), P(c), P(c2), P(cs), S))
    L.check(L.climb_adapter_fused(1, M, d, r, act, P(A), P(wd), P(wu), None, None, P(pre), P(z), P(c), None, P(c2), None, S))
# ---- pair weight-gradient kernel with the bias-gradient MMAs (one tile per pair) and the multi-round fallback ----
for tokens, n_out, k_in in ((1100, 512, 256), (1100, 768, 3072)):
    dy, x = bf(tokens, n_out), bf(tokens, k_in)
    dw, db = torch.zeros(n_out, k_in, device=dev), torch.zeros(n_out, device=dev)
    L.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=True, colsum_a=db)
torch.cuda.synchronize()

SPECS = {"vqa": dict(num_labels=3129, num_images=1, model_type="classification"),
         "vcr": dict(num_labels=4, num_images=1, model_type="multi-choice", num_choices=4)}
HP = {"lr": 1e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8}


def text(n, T=8):
    return {"input_ids": torch.randint(1, 200, (n, T), generator=g).to(dev), "attention_mask": torch.ones(n, T, dtype=torch.long, device=dev),
            "token_type_ids": torch.zeros(n, T, dtype=torch.long, device=dev)}


def learner(**kw):
    cfg = B200ViltConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32,
                         patch_size=16, vocab_size=200, max_position_embeddings=8, **kw)
    torch.manual_seed(0)
    return B200ViltContinualLearner(list(SPECS), B200ViltEncoderWrapper(None, B200ViltModel(cfg), dev), 128, SPECS).to(dev).train()


pm = torch.zeros(2, 64, 80, dtype=torch.long)
pm[0, :48, :80] = 1
pm[1, :64, :48] = 1
# ---- dropout p > 0 inside the ViLT encoder ----
m = learner(hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
opt = m.create_optimizer(HP)
_, lg = m.forward_tensors("vqa", dict(text(3), pixel_values=torch.rand(3, 3, 48, 64, generator=g).to(dev)))
ops.vqa_loss(lg, torch.rand(3, 3129, device=dev).round()).backward()
opt.step(); opt.zero_grad(set_to_none=True)
# ---- VCR: image shared by four choices, fixed and padded ----
m = learner()
opt = m.create_optimizer(HP)
for px, mask in ((torch.rand(2, 3, 48, 64, generator=g), None), (torch.rand(2, 3, 64, 80, generator=g), pm)):
    enc = dict(text(8), pixel_values=px.to(dev))
    if mask is not None:
        enc["pixel_mask"] = mask
    _, lg = m.forward_tensors("vcr", enc)
    ops.cross_entropy_loss(lg, torch.tensor([1, 3], device=dev)).backward()
    opt.step(); opt.zero_grad(set_to_none=True)
# ---- max_image_length: random patch subsets (patch_select) ----
m = learner(max_image_length=5)
opt = m.create_optimizer(HP)
for task, n_text, tgt in (("vqa", 2, None), ("vcr", 8, torch.tensor([0, 2], device=dev))):
    _, lg = m.forward_tensors(task, dict(text(n_text), pixel_values=torch.rand(2, 3, 64, 80, generator=g).to(dev), pixel_mask=pm))
    (ops.vqa_loss(lg, torch.rand(2, 3129, device=dev).round()) if tgt is None else ops.cross_entropy_loss(lg, tgt)).backward()
    opt.step(); opt.zero_grad(set_to_none=True)
# ---- bf16x3 precision engine ----
with ops.precision("bf16x3"):
    m = learner()
    _, lg = m.forward_tensors("vqa", dict(text(2), pixel_values=torch.rand(2, 3, 48, 64, generator=g).to(dev)))
    ops.vqa_loss(lg, torch.rand(2, 3129, device=dev).round()).backward()
# ---- adapters through the fused kernel inside the engine (reduction factor 8 -> r = 16) ----
from climb_b200.modeling import AdapterSpec  # noqa: E402
m = learner()
spec = AdapterSpec.from_config("houlsby")
spec.reduction_factor = 8
m.add_adapter("vcr", spec); m.train_adapter("vcr"); m.set_active_adapters("vcr")
opt = m.create_optimizer(HP)
_, lg = m.forward_tensors("vcr", dict(text(8), pixel_values=torch.rand(2, 3, 48, 64, generator=g).to(dev)))
ops.cross_entropy_loss(lg, torch.tensor([1, 3], device=dev)).backward()
opt.step(); opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("sanitize_r2: done")
