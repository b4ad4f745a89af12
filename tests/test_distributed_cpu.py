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
