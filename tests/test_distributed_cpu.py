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


def test_layer_chunks_and_deferred_optimizer_plan():
    """Host logic of the overlapped exchange: the bottom chunk of the chunked backward is cut once more, and ArenaAdamW's
    chunk table is re-sorted by the completion order of the pending all-reduces."""
    from climb_b200 import distributed as cdist
    sync = cdist.GradSync.__new__(cdist.GradSync)
    sync.layers_per_chunk = 3
    assert sync.layer_chunks(12) == [(11, 9), (8, 6), (5, 3), (2, 1), (0, 0)]
    sync.layers_per_chunk = 2
    assert sync.layer_chunks(4) == [(3, 2), (1, 1), (0, 0)]
    sync.layers_per_chunk = 1
    assert sync.layer_chunks(3) == [(2, 2), (1, 1), (0, 0)]
    # chunks (start, length, group); ranges in ISSUE order: top span first (two buckets), then the bottom span
    chunks = [(0, 100, 0), (100, 100, 1), (200, 100, 0), (300, 100, 0), (400, 50, 1), (1000, 10, 0)]
    ranges = [(250, 350), (350, 450), (0, 250)]
    order, n_free, bounds = cdist.pending_segments(chunks, ranges)
    assert n_free == 1 and order[0] == 5                      # the chunk no range covers goes first, without waiting
    # chunk 3 (300..400) straddles buckets 0 and 1 -> ready after bucket 1; chunk 2 (200..300) straddles bucket 0 and the
    # bottom span -> ready after the bottom span
    assert [chunks[i][0] for i in order] == [1000, 300, 400, 0, 100, 200]
    assert bounds == [1, 3, 6]


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
        deferred = mode == "deferred"
        sync = cdist.attach(learner, bucket_mb=0.05, layers_per_chunk=1, defer_to_optimizer=deferred)   # small buckets: several per chunk
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
        if deferred:
            # finish() handed the in-flight reductions to the optimizer instead of waiting: ranges in issue order (top layer
            # first), each with its work; with gloo's SUM-then-divide stand-in the division is still pending
            pend = arena._pending_reductions
            ok_def = pend is not None and len(pend[0]) > n_layers and all(hi > lo for (lo, hi), _ in pend[0])
            starts = [lo for (lo, hi), _ in pend[0]]
            ok_def = ok_def and starts[0] > starts[-1]                       # top of the arena first, embeddings last
            covered = sum(hi - lo for (lo, hi), _ in pend[0])
            from climb_b200.distributed import pending_segments
            chunks = [(off[n], num[n], 0) for n, p in arena.named_items() if p.requires_grad]
            order, n_free, bounds = pending_segments(chunks, [r for r, _ in pend[0]])
            ok_def = ok_def and n_free == 0 and bounds[-1] == len(chunks) and sorted(order) == list(range(len(chunks)))
            cdist.wait_pending(arena)                                        # what ArenaAdamW.step / prepare_grads / EWC do
            ok_def = ok_def and arena._pending_reductions is None and covered >= sum(num[n] for n, p in arena.named_items() if p.requires_grad)
        else:
            ok_def = getattr(arena, "_pending_reductions", None) is None
        # every chunk top-down, one layer each, tail with the first chunk, embeddings last and alone
        want = [(l, l, _lib.BWD_TAIL if l == n_layers - 1 else 0) for l in range(n_layers - 1, -1, -1)] + [(-1, 0, _lib.BWD_EMBED)]
        ok = calls == want and ok_def
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


@pytest.mark.parametrize("mode", ["full", "frozen_bottom", "adapters", "deferred"])
def test_chunked_backward_overlapped_allreduce_world2_gloo(mode):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_chunked_worker, args=(2, port, out, mode), nprocs=2, join=True)
    assert out[0] and out[1]


# ---------------------------------------------------------------------------------------------------------
# the UNCHANGED harness under torchrun: the learner's training forward works on this rank's rows and returns the gathered outputs
# ---------------------------------------------------------------------------------------------------------
def _harness_worker(rank, world, port, out):
    """climb_b200.distributed.sharded_forward (what B200ViltContinualLearner.forward does in training mode once
    attach_if_distributed ran) around the CPU oracle learner, driven by the restated trainers' train_step. Every rank is handed
    the SAME 5-sample batch (plain loaders, same seed) -- an uneven split, 3 + 2 rows. The trainer must see whole-batch logits
    and compute the whole-batch loss on every rank (whatever it does with its labels: NLVR2Trainer reads them BEFORE it calls
    the model, train_nlvr2.py:128-131), and after the plain mean over ranks every rank must hold the whole-batch gradient."""
    from climb_b200 import distributed as cdist
    from oracle import trainer_oracle as to
    from oracle import vilt_oracle as vo
    from tests.golden_util import ALL_TASKS, TINY, TINY_HW, TINY_T
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        ok, worst_all = True, 0.0
        for task in ("vqa", "nlvr2", "vcr"):
            pool = to.TaskPool(task, 5, TINY, TINY_T, TINY_HW, seed=900)
            items = pool.items(0, 5)
            sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=900, **(dict(layer_scale=6.0, head_scale=20.0) if task == "vcr" else {}))
            proc = to.PoolProcessor({task: pool}, torch.device("cpu"))
            calls = []

            def step(sharded: bool):
                model = to.OracleLearner(TINY, ALL_TASKS, sd, proc)
                inner = model.__call__

                class Sharded:                       # the learner's forward under attach_if_distributed
                    def __call__(self, task_key, images, texts):
                        calls.append(len(texts))
                        if not sharded:
                            return inner(task_key, images, texts)
                        return cdist.sharded_forward(lambda im, tx: inner(task_key, im, tx), images, texts)

                    def __getattr__(self, name):
                        return getattr(model, name)

                batch = to.collate(items)
                logits_fn = Sharded()
                pooled, logits = logits_fn(task, batch["images"], batch["raw_texts"])
                target = batch["target_scores"] if task == "vqa" else batch["labels"]
                loss = vo.task_loss(task, logits, target)          # the WHOLE batch's labels, untouched
                loss.backward()
                return loss, logits, {n: p.grad for n, p in model.named_parameters() if p.grad is not None}

            loss_full, logits_full, g_full = step(False)
            loss_rank, logits_rank, g_rank = step(True)
            ok = ok and logits_rank.shape == logits_full.shape and torch.allclose(logits_rank, logits_full, atol=1e-6)
            ok = ok and abs(loss_rank.item() - loss_full.item()) < 1e-6 * max(1.0, abs(loss_full.item()))
            worst = 0.0
            gscale = max(g.norm().item() for g in g_full.values())     # (the key-bias gradients are analytically zero: fp32 noise)
            ok = ok and set(g_rank) == set(g_full)
            for n, g in g_rank.items():
                flat = g.clone().flatten()
                cdist.allreduce_mean_(flat, [(0, flat.numel())])        # GradSync's plain mean over ranks
                ref = g_full[n].flatten()
                worst = max(worst, (flat - ref).norm().item() / max(ref.norm().item(), 1e-3 * gscale))
            ok = ok and worst < 2e-5
            worst_all = max(worst_all, worst)
        # a batch smaller than the world leaves a rank without rows: loud error, not a hang
        try:
            cdist.sharded_forward(lambda im, tx: None, [0], ["a"])
            ok = ok and rank == 0 and False
        except RuntimeError:
            ok = ok and rank == 1
        except Exception:
            ok = ok and rank == 0          # rank 0 has the row and fails later on the dummy forward
        out[rank] = (bool(ok), worst_all)
    finally:
        dist.destroy_process_group()


def test_sharded_forward_gives_the_unchanged_harness_whole_batch_outputs_and_gradients_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_harness_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0], dict(out)


def test_learner_train_eval_drive_the_sharding_mode_and_single_process_is_untouched():
    from climb_b200 import distributed as cdist
    from climb_b200.modeling import model_configs
    from tests.test_host_logic import _learner
    learner = _learner()
    learner.eval()
    assert cdist._training_mode is False
    learner.train()
    assert cdist._training_mode is True
    batch = {"images": [1, 2, 3], "raw_texts": ["a", "b", "c"], "labels": torch.arange(3)}
    inputs = model_configs["vilt-b200"]["batch2inputs_converter"](batch)       # the reference's converter: the batch dict is never edited
    assert inputs == {"images": [1, 2, 3], "texts": ["a", "b", "c"]} and len(batch["labels"]) == 3
    assert cdist.attach_if_distributed(learner) is None and not getattr(learner, "_ddp_forward_shard", False)
