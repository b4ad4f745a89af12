"""The persistent attention kernels for compute-sanitizer (dev tool):
    compute-sanitizer --tool memcheck python tools/sanitize_attn.py
B x H = 14 x 12 = 168 items > 148 SMs: some CTAs walk two items (the rolling operand prefetch, the dQ / dK / dV drains and the
next-item scalars all run), L = 237 (two key / query tiles with a ragged tail), key mask on, bias-gradient sums on. The results
are compared with the one-tile-per-CTA kernels (CLIMB_ATTN_V1=1 in a second process is not needed: a torch reference is)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import _lib as L  # noqa: E402

B, Lq, H = int(os.environ.get("B", 14)), 237, 12
torch.manual_seed(0)
qkv = torch.randn(B, Lq, 3 * H * 64, device="cuda").bfloat16()
lens = torch.randint(Lq // 2, Lq + 1, (B,), device="cuda")
kb = (1.0 - (torch.arange(Lq, device="cuda")[None, :] < lens[:, None]).float()) * -10000.0
scale = 1.0 / math.sqrt(64)
ctx, lse = L.attention_fwd(qkv, kb, B, Lq, H, scale)
dctx = torch.randn(B, Lq, H * 64, device="cuda").bfloat16()
cs = torch.zeros(3 * H * 64, device="cuda")
dqkv = L.attention_bwd(qkv, kb, ctx, dctx, lse, B, Lq, H, scale, colsum=cs)
torch.cuda.synchronize()
q, k, v = qkv.float().view(B, Lq, 3, H, 64).permute(2, 0, 3, 1, 4)
p = torch.softmax(q @ k.transpose(-1, -2) * scale + kb[:, None, None, :], -1)
ref = (p @ v).permute(0, 2, 1, 3).reshape(B, Lq, H * 64)
err = ((ctx.float() - ref).norm() / ref.norm()).item()
print(f"sanitize_attn: B={B} ctx rel err {err:.2e} dqkv finite {bool(torch.isfinite(dqkv.float()).all())} colsum finite {bool(torch.isfinite(cs).all())}")
assert err < 6e-3
