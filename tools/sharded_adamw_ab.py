"""A/B for SURVEY 8 f-1's collective half (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_adamw_ab.py

  A  product path     gradient ALL-REDUCE (mean) of every span from inside the chunked backward, AdamW on the whole arena
                      on every rank (climb_b200.distributed.GradSync + ArenaAdamW)
  B  sharded          gradient REDUCE-SCATTER of the same spans from inside the chunked backward (half the bytes on the wire
                      during the backward), AdamW only on this rank's 1/N of every bucket, then ALL-GATHER of the updated
                      fp32 parameters and of their bf16 shadow (6 bytes per parameter; nothing left to hide it behind)

Same model / batch / step as bench.py (ViLT-base, B = 64 per GPU, fwd + BCE x 3129 + bwd + AdamW), CUDA events, max over
ranks. B is a measurement harness around the library's own kernels (climb_adamw_step with a chunk table cut to the owned
ranges), not a product path: the table it prints is the evidence for keeping A."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from climb_b200 import _lib, ops  # noqa: E402
from climb_b200 import distributed as cdist  # noqa: E402
from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel  # noqa: E402
from climb_b200.optim import _upload_chunks  # noqa: E402

import ctypes  # noqa: E402


class ShardedSync(cdist.GradSync):
    """Same chunked-backward protocol as GradSync, but every bucket is reduce-scattered: rank r ends up with the mean of
    elements [bs + r m, bs + (r + 1) m) of bucket [bs, be), m = (be - bs) // world (the < world leftover elements of a
    bucket are all-reduced and updated redundantly)."""

    def __init__(self, learner, **kw):
        super().__init__(learner, **kw)
        self.owned, self.buckets = [], []

    def begin(self, arena):
        super().begin(arena)
        self.owned, self.buckets = [], []

    def reduce_range(self, arena, lo, hi):
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        for s, e in self._spans(arena):
            s2, e2 = max(s, lo), min(e, hi)
            if e2 <= s2:
                continue
            for bs, be in cdist.bucketize([(s2, e2)], self.bucket_elems):
                m = (be - bs) // world
                if m > 0:
                    full = arena.grad[bs: bs + m * world]
                    self._works.append(dist.reduce_scatter_tensor(full[rank * m: (rank + 1) * m], full, op=dist.ReduceOp.AVG,
                                                                  group=self.group, async_op=True))
                    self.owned.append((bs + rank * m, bs + (rank + 1) * m))
                    self.buckets.append((bs, m))
                if bs + m * world < be:
                    self._works.append(dist.all_reduce(arena.grad[bs + m * world: be], op=dist.ReduceOp.AVG, group=self.group,
                                                       async_op=True))
                    self.owned.append((bs + m * world, be))

    def finish(self, arena):
        self._wait_loose()
        for w in self._works:
            w.wait()
        self._works = []


class ShardedAdamW:
    """AdamW over the owned ranges of the arena + the all-gather of what it updated; the task heads (replicated,
    all-reduced) go through the ordinary ArenaAdamW."""

    def __init__(self, learner, sync: ShardedSync, hp):
        self.inner = learner.create_optimizer(hp)
        self.sync, self.learner = sync, learner
        self.arena = learner.get_encoder().vilt._arena
        self.arena.sync()
        self.m = torch.zeros_like(self.arena.theta)
        self.v = torch.zeros_like(self.arena.theta)
        self.step_n = 0
        self.table = None
        self.arena_params = {id(p) for _, p in self.arena.named_items()}

    def _build(self):
        a = self.arena
        chunks = []
        owned = sorted(self.sync.owned)
        for gi, g in enumerate(self.inner.param_groups):
            for p in g["params"]:
                if id(p) not in self.arena_params or p.grad is None:
                    continue
                s = (p.data_ptr() - a.theta.data_ptr()) // 4
                e = s + p.numel()
                for lo, hi in owned:
                    x0, x1 = max(s, lo), min(e, hi)
                    o = x0
                    while o < x1:
                        n = min(1 << 16, x1 - o)
                        chunks.append((o, n, gi))
                        o += n
        self.table = (_upload_chunks(chunks, a.theta.device), len(chunks), tuple(owned))

    @torch.no_grad()
    def step(self):
        a = self.arena
        if self.table is None or self.table[2] != tuple(sorted(self.sync.owned)):
            self._build()
        g0 = self.inner.param_groups
        lr = (ctypes.c_float * len(g0))(*[float(g["lr"]) for g in g0])
        wd = (ctypes.c_float * len(g0))(*[float(g["weight_decay"]) for g in g0])
        self.step_n += 1
        b1, b2 = g0[0]["betas"]
        a.refresh_shadow()
        _lib.check(_lib.climb_adamw_step(_lib.ptr(a.theta), _lib.ptr(a.grad), _lib.ptr(self.m), _lib.ptr(self.v), _lib.ptr(a.shadow),
                                         _lib.ptr(self.table[0]), self.table[1], lr, wd, len(g0), b1, b2, g0[0]["eps"], self.step_n,
                                         _lib.stream()))
        world, rank = dist.get_world_size(), dist.get_rank()
        works = []
        with dist._coalescing_manager(async_ops=True) as cm:
            for bs, m in self.sync.buckets:
                full = a.theta[bs: bs + m * world]
                dist.all_gather_into_tensor(full, full[rank * m: (rank + 1) * m])
        works.append(cm)
        with dist._coalescing_manager(async_ops=True) as cm2:
            for bs, m in self.sync.buckets:
                full = a.shadow[bs: bs + m * world]
                dist.all_gather_into_tensor(full, full[rank * m: (rank + 1) * m])
        works.append(cm2)
        for w in works:
            w.wait()
        # the heads: replicated AdamW on their (all-reduced) gradients
        saved = {}
        for _, p in a.named_items():
            if p.grad is not None:
                saved[p] = p.grad
                p.grad = None
        self.inner.step()
        for p, g in saved.items():
            p.grad = g

    def zero_grad(self, set_to_none=True):
        self.inner.zero_grad(set_to_none=set_to_none)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    specs = {"vqa": dict(num_labels=bench.N_LABELS, num_images=1, model_type="classification")}
    hp = {"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8}
    batches = [{k: v.to(dev) for k, v in bench.make_host_batch(args.batch, 1000 * rank + i, pin=False).items()} for i in range(4)]
    rows = []
    for variant in ("all-reduce + replicated AdamW (product path)", "all-reduce, AdamW deferred span by span (product path, bench default)",
                    "reduce-scatter + sharded AdamW + all-gather"):
        torch.manual_seed(42)
        learner = B200ViltContinualLearner(["vqa"], B200ViltEncoderWrapper(None, B200ViltModel(B200ViltConfig()), dev), 768, specs).to(dev)
        learner.train()
        if variant.startswith("reduce-scatter"):
            sync = ShardedSync(learner)
            opt = ShardedAdamW(learner, sync, hp)
        else:
            sync = cdist.attach(learner, defer_to_optimizer="deferred" in variant)
            opt = learner.create_optimizer(hp)

        def step(b):
            _, logits = learner.forward_tensors("vqa", {k: b[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values")})
            loss = ops.vqa_loss(logits, b["target"])
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss

        for i in range(args.warmup):
            step(batches[i % 4])
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = step(batches[i % 4])
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # replicas must still agree after the run
        flat = torch.cat([p.detach().flatten()[:4096] for p in learner.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1 if torch.equal(flat, ref) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        rows.append({"variant": variant, "n_gpus": world, "ms_per_step": round(float(t.item()), 3),
                     "samples_per_s": round(world * args.batch / float(t.item()) * 1e3, 1), "last_loss": round(float(loss.detach()), 4),
                     "replicas_identical": bool(int(same.item()))})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
        sync.detach()
        del learner, opt, sync
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(rows, open(f"gpurun_out/sharded_adamw_ab_{world}gpu.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
