"""Steady-state step time of the bench workload (MODE=vqa) or the adapter configuration (MODE=adapters: NLVR2 pairs, Houlsby
rf 16, base frozen = BASELINE config 3 on one GPU), CUDA events over 20 steps after warm-up. Dev tool."""
import os
import runpy
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("WARM", "5")
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "one_step.py"))
step = ns["step"]
for _ in range(3):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"MODE={os.environ.get('MODE', 'vqa')} B={ns['B']} step {ms:.3f} ms  ({ns['B'] / ms * 1e3:.0f} sequences/s)")
