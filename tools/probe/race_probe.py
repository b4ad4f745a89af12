"""racecheck bisect (dev): WHICH=wgrad_cs | wgrad | pair | adapter"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from climb_b200 import _lib as L
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
bf = lambda *s: (torch.randn(*s, generator=g) * 0.1).to(dev).bfloat16()
which = os.environ.get("WHICH", "wgrad_cs")
if which in ("wgrad_cs", "wgrad"):
    dy, x = bf(1100, 512), bf(1100, 256)
    dw, db = torch.zeros(512, 256, device=dev), torch.zeros(512, device=dev)
    L.gemm(dy, x, dw, a_mn_major=True, b_mn_major=True, accumulate=True, colsum_a=db if which == "wgrad_cs" else None)
elif which == "pair":
    a, b = bf(2500, 192), bf(2304, 192)
    ob = torch.empty(2500, 2304, device=dev, dtype=torch.bfloat16)
    L.gemm(a, b, ob)
else:
    M, d, r = 300, 128, 16
    A, wd, wu = bf(M, d), bf(r, d), bf(d, r)
    bd, bu = torch.randn(r, device=dev), torch.randn(d, device=dev)
    pre, z = torch.empty(M, r, device=dev, dtype=torch.bfloat16), torch.empty(M, r, device=dev, dtype=torch.bfloat16)
    c, c2 = torch.randn(M, d, device=dev), torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    L.check(L.climb_adapter_fused(0, M, d, r, L.EPI_SWISH, L.ptr(A), L.ptr(wd), L.ptr(wu), L.ptr(bd), L.ptr(bu), L.ptr(pre), L.ptr(z), L.ptr(c), L.ptr(c), L.ptr(c2), None, L.stream()))
torch.cuda.synchronize()
print("done", which)
