"""Minimal driver for an ncu capture of the CTA-pair GEMM kernels (no model, two launches):
    ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 2 -c 2 -f -o gpurun_out/prof_pair python tools/ncu_pair.py
Launch 1: FC2 weight gradient at B = 64 (gemm_pair_wgrad_kernel); launch 2: QKV forward (gemm_pair_kernel<FK_BF16>)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import _lib as L  # noqa: E402

M, d, ff = 64 * 237, 768, 3072
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.05).bfloat16()
dy, inter, gw2 = bf(M, d), bf(M, ff), torch.zeros(d, ff, device="cuda")
x, w_qkv, o_qkv, bias = bf(M, d), bf(3 * d, d), torch.empty(M, 3 * d, device="cuda", dtype=torch.bfloat16), torch.randn(3 * d, device="cuda")
for _ in range(2):          # first pass = warm-up (skipped with -s 2), second pass = the captured launches
    L.gemm(dy, inter, gw2, a_mn_major=True, b_mn_major=True, accumulate=True, M=d, N=ff, K=M)
    L.gemm(x, w_qkv, o_qkv, bias=bias)
torch.cuda.synchronize()
print("ok")
