"""Experiment (dev): AdamW of the layers whose gradients are complete launched on a SIDE stream while the backward of the layers
below is still running, on ONE GPU (the data-parallel path already does this behind the all-reduces). Measures the step time
against the plain loop on the same box. Not a product path: it only answers 'would hiding AdamW behind the backward pay?'."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from climb_b200 import _lib, ops  # noqa: E402
from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel  # noqa: E402
from climb_b200.optim import _upload_chunks  # noqa: E402

dev = torch.device("cuda")
B = 64
specs = {"vqa": dict(num_labels=3129, num_images=1, model_type="classification")}
torch.manual_seed(42)
learner = B200ViltContinualLearner(["vqa"], B200ViltEncoderWrapper(None, B200ViltModel(B200ViltConfig()), dev), 768, specs).to(dev).train()
opt = learner.create_optimizer({"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8})
batch = {k: v.to(dev) for k, v in bench.make_host_batch(B, 0, pin=False).items()}
vilt = learner.get_encoder().vilt
CHUNK = int(os.environ.get("CHUNK", 3))


class Overlap:
    """GradSync-shaped hook object: the chunked backward calls begin / reduce_range / finish."""
    layers_per_chunk = CHUNK

    def __init__(self):
        self.side = torch.cuda.Stream()
        self.m = self.v = None
        self.step = 0
        self.tables = {}
        self.done = []

    def layer_chunks(self, n):
        return [(f, max(0, f - CHUNK + 1)) for f in range(n - 1, -1, -CHUNK)]

    def begin(self, arena):
        if self.m is None:
            self.m, self.v = torch.zeros_like(arena.theta), torch.zeros_like(arena.theta)
        self.step += 1
        self.done = []

    def _table(self, arena, lo, hi):
        key = (lo, hi)
        if key not in self.tables:
            chunks = []
            for gi, g in enumerate(opt.param_groups):
                for p in g["params"]:
                    s = (p.data_ptr() - arena.theta.data_ptr()) // 4
                    if not (0 <= s < arena.size) or not p.requires_grad:
                        continue
                    if lo <= s < hi:
                        for o in range(0, p.numel(), 1 << 16):
                            chunks.append((s + o, min(1 << 16, p.numel() - o), gi))
            self.tables[key] = (_upload_chunks(chunks, dev), len(chunks)) if chunks else (None, 0)
        return self.tables[key]

    def _adamw(self, arena, lo, hi, stream):
        table, n = self._table(arena, lo, hi)
        if n == 0:
            return
        g0 = opt.param_groups
        lr = (ctypes.c_float * len(g0))(*[float(g["lr"]) for g in g0])
        wd = (ctypes.c_float * len(g0))(*[float(g["weight_decay"]) for g in g0])
        b1, b2 = g0[0]["betas"]
        _lib.check(_lib.climb_adamw_step(_lib.ptr(arena.theta), _lib.ptr(arena.grad), _lib.ptr(self.m), _lib.ptr(self.v), _lib.ptr(arena.shadow),
                                         _lib.ptr(table), n, lr, wd, len(g0), b1, b2, g0[0]["eps"], self.step, stream))

    def reduce_range(self, arena, lo, hi):
        ev = torch.cuda.Event()
        ev.record()
        self.side.wait_event(ev)
        self._adamw(arena, lo, hi, ctypes.c_void_p(self.side.cuda_stream))
        self.done.append((lo, hi))

    def finish(self, arena):
        pass


ov = Overlap()
MODE = os.environ.get("OVERLAP", "1") == "1"
if MODE:
    vilt.grad_sync = ov


def step():
    enc = {k: batch[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values")}
    _, logits = learner.forward_tensors("vqa", enc)
    loss = ops.vqa_loss(logits, batch["target"])
    loss.backward()
    if MODE:
        # what the chunked backward did not cover is nothing here (the last reduce_range call runs down to offset 0); the heads
        # go through the ordinary optimizer: hide the arena's parameters from it
        torch.cuda.current_stream().wait_stream(ov.side)
        saved = {}
        for _, p in vilt._arena.named_items():
            if p.grad is not None:
                saved[p] = p.grad
                p.grad = None
        opt.step()
    else:
        opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


for _ in range(6):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    loss = step()
b.record()
torch.cuda.synchronize()
print(f"OVERLAP={int(MODE)} CHUNK={CHUNK}: step {a.elapsed_time(b) / 20:.3f} ms, loss {float(loss.detach()):.4f}")
