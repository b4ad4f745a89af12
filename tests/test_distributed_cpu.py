"""world_size-2 gloo tests (CPU) of the data-parallel host logic: span / bucket computation over the
flat arena, the bucketed mean all-reduce and batch sharding. The NCCL path runs the same functions on
CUDA tensors (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from climb_b200.distributed import allreduce_mean_, bucketize, shard_batch, trainable_spans


def test_spans_and_buckets():
    offsets = {"a": 0, "b": 128, "c": 256, "d": 1024}
    numels = {"a": 100, "b": 128, "c": 700, "d": 10}
    order = ["a", "b", "c", "d"]
    assert trainable_spans(offsets, numels, order, {"a", "b", "c", "d"}) == [(0, 956), (1024, 1034)]
    assert trainable_spans(offsets, numels, order, {"a", "c"}) == [(0, 100), (256, 956)]
    assert trainable_spans(offsets, numels, order, set()) == []
    assert bucketize([(0, 956), (1024, 1034)], 400) == [(0, 400), (400, 800), (800, 956), (1024, 1034)]


def test_shard_batch_keeps_groups_together():
    batch = {"x": torch.arange(16), "names": list("abcdefghijklmnop"), "k": "vcr"}
    parts = [shard_batch(batch, r, 2, group_size=4) for r in range(2)]
    assert torch.equal(torch.cat([p["x"] for p in parts]), batch["x"])
    assert all(len(p["x"]) % 4 == 0 for p in parts)
    assert parts[0]["names"] + parts[1]["names"] == batch["names"] and parts[0]["k"] == "vcr"
    three = [shard_batch({"x": torch.arange(10)}, r, 3) for r in range(3)]
    assert sum(len(p["x"]) for p in three) == 10


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    untouched = flat.clone()
    buckets = bucketize([(0, 300), (512, 900)], 128)
    allreduce_mean_(flat, buckets)
    expect = torch.arange(1000, dtype=torch.float32) * 1.5          # mean of x1 and x2
    ok = torch.allclose(flat[0:300], expect[0:300]) and torch.allclose(flat[512:900], expect[512:900])
    ok = ok and torch.equal(flat[300:512], untouched[300:512]) and torch.equal(flat[900:], untouched[900:])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucketed_mean_allreduce_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


# ---------------------------------------------------------------------------------------------------------
# the overlapped (chunked) backward + GradSync.begin / reduce_range / finish, as bench.py --gpus N runs it
# ---------------------------------------------------------------------------------------------------------
def _chunked_worker(rank, world, port, out, mode):
    """The REAL B200ViltModel._run_backward and GradSync on CPU tensors over gloo. Stubbed for the test: the C entry
    point (records its (first, last, parts) arguments and writes rank-dependent 'gradients' into the spans that call
    completes), _lib.ptr / stream (no device here) and ReduceOp.AVG (gloo has no AVG: SUM then divide)."""
    from climb_b200 import _lib
    from climb_b200 import distributed as cdist
    from climb_b200.modeling import vilt_model as vm
    from tests.test_host_logic import _learner
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        learner = _learner()
        if mode == "frozen_bottom":
            learner.get_encoder().freeze_bottom_k_layers(1)
        elif mode == "adapters":                 # base frozen, one Houlsby adapter trains: two small spans per layer
            learner.add_adapter("vqa", "houlsby")
            learner.train_adapter("vqa")
        vilt = learner.get_encoder().vilt
        arena = vilt._arena
        arena.sync(allow_cpu=True)
        off, num = arena.offsets, arena.numels
        n_layers = len(vilt.encoder.layer)
        sync = cdist.attach(learner, bucket_mb=0.05, layers_per_chunk=1)        # small buckets: several per chunk
        calls, reduced = [], [0]
        pattern = torch.arange(arena.size, dtype=torch.float32) % 97 + 1

        def fake_backward(dims, params, batch, theta, shadow, ws, ws_bytes, scratch, scratch_n, dpooled, grad, first, last,
                          parts, stream):
            calls.append((first, last, parts))
            names = [n for n in off if (n.startswith("encoder.layer.") and first >= 0 and last <= int(n.split(".")[2]) <= first)
                     or (parts & _lib.BWD_TAIL and n.startswith(("layernorm.", "pooler.")))
                     or (parts & _lib.BWD_EMBED and n.startswith("embeddings."))]
            rg = dict(arena.named_items())
            for n in names:
                if rg[n].requires_grad:
                    arena.grad[off[n]: off[n] + num[n]] += (rank + 1) * pattern[off[n]: off[n] + num[n]]
            return 0

        real_all_reduce = dist.all_reduce

        class Done:
            def wait(self):
                return True

        def all_reduce(t, op=dist.ReduceOp.SUM, group=None, async_op=False):
            reduced[0] += t.numel()
            if op == dist.ReduceOp.AVG:
                real_all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                t.div_(world)
            else:
                real_all_reduce(t, op=op, group=group)
            return Done() if async_op else None

        _lib.climb_vilt_backward, _lib.climb_vilt_backward_scratch_bytes = fake_backward, (lambda *a: 0)
        _lib.ptr, _lib.stream = (lambda t: 0 if t is None else t.data_ptr()), (lambda: 0)
        dist.all_reduce = all_reduce
        st = vilt._static_tables()
        call = vm._Call()
        call.dims, call.params, call.layers, call.batch = st["dims"], st["params"], st["layers"], _lib.ViltBatchC()
        call.trainable, call.workspace, call.ws_bytes, call.arena_theta, call.keep = st["trainable"], None, 0, arena.theta, []
        vilt._run_backward(call, torch.zeros(2, vilt.config.hidden_size))
        # every chunk top-down, one layer each, tail with the first chunk, embeddings last and alone
        want = [(l, l, _lib.BWD_TAIL if l == n_layers - 1 else 0) for l in range(n_layers - 1, -1, -1)] + [(-1, 0, _lib.BWD_EMBED)]
        ok = calls == want
        mean = (1 + world) / 2.0
        n_trainable = 0
        for n, p in arena.named_items():
            g = arena.grad[off[n]: off[n] + num[n]]
            if p.requires_grad:
                n_trainable += num[n]
                ok = ok and torch.allclose(g, mean * pattern[off[n]: off[n] + num[n]]) and p.grad is not None \
                    and p.grad.data_ptr() == g.data_ptr()
            else:
                ok = ok and bool((g == 0).all()) and p.grad is None
        # every trainable element went through exactly one all-reduce (alignment padding between neighbours may ride along)
        ok = ok and n_trainable <= reduced[0] <= n_trainable + 64 * len(off)
        # the task heads live outside the arena: reduced from post-accumulate hooks
        head = learner.task_layer["vqa"][0].weight
        head.grad = None
        (head * (rank + 1)).sum().backward()
        ok = ok and torch.allclose(head.grad, torch.full_like(head, mean))
        sync.detach()
        ok = ok and vilt.grad_sync is None
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["full", "frozen_bottom", "adapters"])
def test_chunked_backward_overlapped_allreduce_world2_gloo(mode):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_chunked_worker, args=(2, port, out, mode), nprocs=2, join=True)
    assert out[0] and out[1]
