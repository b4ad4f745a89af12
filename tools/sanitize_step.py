"""Tiny forward + backward + optimizer steps of every engine path, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_step.py
Covers: ViLT learner (fixed resolution, padded / ragged batch, adapters), ViLT-BERT (train mode: both dropouts live),
EWC penalty + Fisher, losses, AdamW, the long-sequence (L > 256) attention path, the big-tile GEMM kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import _lib as L, ops  # noqa: E402
from climb_b200.modeling import (B200BertConfig, B200BertModel, B200ViltBertContinualLearner, B200ViltBertEncoderWrapper,  # noqa: E402
                                 B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel)

dev = torch.device("cuda")
SPECS = {"vqa": dict(num_labels=3129, num_images=1, model_type="classification"),
         "nlvr2": dict(num_labels=2, num_images=2, model_type="classification"),
         "vcr": dict(num_labels=4, num_images=1, model_type="multi-choice", num_choices=4)}
cfg = B200ViltConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32,
                     patch_size=16, vocab_size=200, max_position_embeddings=8)
HP = {"lr": 1e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8}
g = torch.Generator().manual_seed(0)


def text(n, T=8):
    return {"input_ids": torch.randint(1, 200, (n, T), generator=g).to(dev), "attention_mask": torch.ones(n, T, dtype=torch.long, device=dev),
            "token_type_ids": torch.zeros(n, T, dtype=torch.long, device=dev)}


torch.manual_seed(0)
m = B200ViltContinualLearner(list(SPECS), B200ViltEncoderWrapper(None, B200ViltModel(cfg), dev), 128, SPECS).to(dev).train()
opt = m.create_optimizer(HP)
# fixed resolution, VQA
enc = dict(text(3), pixel_values=torch.rand(3, 3, 48, 64, generator=g).to(dev))
_, lg = m.forward_tensors("vqa", enc)
ops.vqa_loss(lg, torch.rand(3, 3129, device=dev).round()).backward()
opt.step(); opt.zero_grad(set_to_none=True)
# padded batch (ragged), NLVR2 pairs, mask on the GPU and on the host
pm = torch.zeros(4, 64, 80, dtype=torch.long)
for k, (h, w) in enumerate([(64, 80), (48, 48), (32, 64), (16, 80)]):
    pm[k, :h, :w] = 1
for mask in (pm.to(dev), pm):
    enc = dict(text(2), pixel_values=torch.rand(4, 3, 64, 80, generator=g).to(dev), pixel_mask=mask)
    _, lg = m.forward_tensors("nlvr2", enc)
    ops.cross_entropy_loss(lg, torch.tensor([0, 1], device=dev)).backward()
    opt.step(); opt.zero_grad(set_to_none=True)
# long sequence: 8 + 1 + 256 tokens
enc = dict(text(2), pixel_values=torch.rand(2, 3, 256, 256, generator=g).to(dev))
_, lg = m.forward_tensors("vqa", enc)
lg.float().sum().backward()
opt.zero_grad(set_to_none=True)
# adapters
m.add_adapter("vcr", "houlsby"); m.train_adapter("vcr"); m.set_active_adapters("vcr")
aopt = m.create_optimizer(HP)
enc = dict(text(8), pixel_values=torch.rand(2, 3, 48, 64, generator=g).to(dev))
_, lg = m.forward_tensors("vcr", enc)
ops.cross_entropy_loss(lg, torch.tensor([1, 3], device=dev)).backward()
aopt.step(); aopt.zero_grad(set_to_none=True)
# ViLT-BERT in train mode (dropouts live) + EWC-style penalty
bcfg = B200BertConfig(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, max_position_embeddings=16)
vb = B200ViltBertContinualLearner(list(SPECS), B200ViltBertEncoderWrapper(None, B200ViltModel(cfg), B200BertModel(bcfg), dev), 128, SPECS).to(dev).train()
vopt = vb.create_optimizer(HP)
enc = dict(text(8), pixel_values=torch.rand(2, 3, 48, 64, generator=g).to(dev))
_, lg = vb.forward_tensors("vcr", enc)
arena = vb.get_encoder().vilt._arena
pen = ops.ewc_penalty(arena, arena.theta.detach().clone() + 1e-3, torch.rand_like(arena.theta) * 1e-3, 100.0,
                      [(n, p) for n, p in arena.named_items() if "word_embeddings" not in n])
(ops.cross_entropy_loss(lg, torch.tensor([0, 2], device=dev)) + pen).backward()
vopt.step(); vopt.zero_grad(set_to_none=True)
# big-tile GEMM kernels (>= 148 tiles) incl. the specialised epilogues
M, N, K = 2500, 2304, 192
a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16); aux = torch.empty_like(ob); of = torch.empty(M, N, device=dev)
L.gemm(a, b, ob, bias=bias)
L.gemm(a, b, ob, bias=bias, epilogue=L.EPI_GELU_SAVE_GRAD, aux=aux)
L.gemm(a, b, ob, epilogue=L.EPI_MUL_AUX, aux=aux)
L.gemm(a, b, of, bias=bias, residual=res, aux=aux, c2=ob)
torch.cuda.synchronize()
print("sanitize_step: done")
