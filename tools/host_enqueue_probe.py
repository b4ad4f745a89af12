"""How long does the HOST need to enqueue one training step (dev tool)? The GPU is idle when the call starts, so the time of
step() until it returns is pure host work (Python + launches); the device time comes from CUDA events around the same step.
If host >= device the loop is launch-bound and the e2e number (one step of look-ahead) suffers first."""
import os
import runpy
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("WARM", "5")
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "one_step.py"))
step = ns["step"]
host, dev = [], []
for _ in range(8):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    step()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    host.append((t1 - t0) * 1e3)
    dev.append(a.elapsed_time(b))
print(f"host enqueue per step: median {sorted(host)[len(host) // 2]:.2f} ms (min {min(host):.2f}); device per step (cold start each): median {sorted(dev)[len(dev) // 2]:.2f} ms")
import cProfile
import pstats
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
