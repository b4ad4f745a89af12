"""Minimal driver for the round-2 ncu captures (no model; every kernel launched twice, the second pass is captured):
    ncu --set full --clock-control none --import-source on -k regex:'gemm_pair_wgrad|adapter_fused|ewc_penalty|fisher|adamw' \
        -s 7 -c 7 -f -o gpurun_out/prof_r2 python tools/ncu_r2.py
  1  gemm_pair_wgrad_kernel   FC1 weight gradient at B = 64 WITH its bias gradient (colsum_a: the extra MMA against ones)
  2  gemm_pair_wgrad_kernel   FC2 weight gradient + bias gradient
  3  adapter_fused_kernel     forward, M = 15168, d = 768, r = 48 (Houlsby rf 16)
  4  adapter_fused_kernel     backward
  5  ewc_penalty kernel(s)    loss + gradient over a 113 M-parameter arena
  6  fisher_accumulate
  7  adamw_kernel             one arena-wide step incl. the bf16 shadow"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import _lib as L  # noqa: E402
from climb_b200.optim import _upload_chunks  # noqa: E402

M, d, ff, r = 64 * 237, 768, 3072, 48
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.05).bfloat16()
du, h2, gw1, gb1 = bf(M, ff), bf(M, d), torch.zeros(ff, d, device="cuda"), torch.zeros(ff, device="cuda")
dy, inter, gw2, gb2 = bf(M, d), bf(M, ff), torch.zeros(d, ff, device="cuda"), torch.zeros(d, device="cuda")
A, wd, wu = bf(M, d), bf(r, d), bf(d, r)
bd, bu = torch.randn(r, device="cuda"), torch.randn(d, device="cuda")
pre, z = torch.empty(M, r, device="cuda", dtype=torch.bfloat16), torch.empty(M, r, device="cuda", dtype=torch.bfloat16)
c = torch.randn(M, d, device="cuda")
c2 = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
cs = torch.zeros(r, device="cuda")
n = 113_000_000
theta, star, fisher, grad = (torch.randn(n, device="cuda") * 0.02 for _ in range(4))
fisher.abs_()
m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
partials, loss = torch.zeros(2048, device="cuda"), torch.zeros(1, device="cuda")
table = _upload_chunks([(o, min(1 << 16, n - o), 0) for o in range(0, n, 1 << 16)], theta.device)
n_chunks = (n + (1 << 16) - 1) >> 16
lr, wdk = (ctypes.c_float * 1)(1e-4), (ctypes.c_float * 1)(1e-2)
P, S = L.ptr, L.stream()
for _ in range(2):
    L.gemm(du, h2, gw1, a_mn_major=True, b_mn_major=True, accumulate=True, M=ff, N=d, K=M, colsum_a=gb1)
    L.gemm(dy, inter, gw2, a_mn_major=True, b_mn_major=True, accumulate=True, M=d, N=ff, K=M, colsum_a=gb2)
    L.check(L.climb_adapter_fused(0, M, d, r, L.EPI_SWISH, P(A), P(wd), P(wu), P(bd), P(bu), P(pre), P(z), P(c), P(c), P(c2), None, S))
    L.check(L.climb_adapter_fused(1, M, d, r, L.EPI_SWISH, P(A), P(wd), P(wu), None, None, P(pre), P(z), P(c), P(c), P(c2), P(cs), S))
    L.check(L.climb_ewc_penalty(P(theta), P(star), P(fisher), n, 100.0, P(partials), 2048, P(loss), P(grad), 1.0, None, S))
    L.check(L.climb_fisher_accumulate(P(grad), P(fisher), n, S))
    L.check(L.climb_adamw_step(P(theta), P(grad), P(m), P(v), P(shadow), P(table), n_chunks, lr, wdk, 1, 0.9, 0.98, 1e-8, 1, S))
torch.cuda.synchronize()
print("ok")
