"""Data-parallel gradient synchronisation for the B200 ViLT path (the reference is
single-process, so this layer is new). One process per GPU, torch.distributed over NCCL/NVLink.

The path shards on BATCH only: every rank runs the same task on its own rows of the step's batch, the
encoder has no cross-row operation, and the single exchange per step is the gradient all-reduce
(average). Because all encoder gradients live in ONE flat arena (climb_b200/arena.py) the exchange is a
handful of large all-reduces over contiguous spans of that buffer instead of ~230 per-tensor calls:

  * spans = maximal runs of trainable parameters in arena order (full fine-tuning: one 446 MB span;
    adapters: 2 small spans per layer), cut into buckets of at most `bucket_mb`;
  * task-head gradients (ordinary autograd tensors outside the arena) are reduced from
    post-accumulate-grad hooks as soon as autograd produces them, i.e. while the encoder backward
    is still running.

EWC's penalty / Fisher and the optimizer are purely local (replicated parameters): no communication.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def trainable_spans(offsets, numels, names_in_order: Sequence[str], trainable: set, align: int = 64) -> List[Tuple[int, int]]:
    """Maximal contiguous [start, end) element ranges of the arena covered by trainable parameters
    (alignment padding between two trainable neighbours is bridged)."""
    spans: List[Tuple[int, int]] = []
    for n in names_in_order:
        if n not in trainable:
            continue
        s, e = offsets[n], offsets[n] + numels[n]
        if spans and s - spans[-1][1] < align:
            spans[-1] = (spans[-1][0], e)
        else:
            spans.append((s, e))
    return spans


def bucketize(spans: Sequence[Tuple[int, int]], bucket_elems: int) -> List[Tuple[int, int]]:
    out = []
    for s, e in spans:
        while e - s > bucket_elems:
            out.append((s, s + bucket_elems))
            s += bucket_elems
        if e > s:
            out.append((s, e))
    return out


def allreduce_mean_(flat: torch.Tensor, buckets: Sequence[Tuple[int, int]], group=None) -> None:
    """In-place average of the listed ranges of `flat` over the process group. Buckets are issued
    back to front: the backward pass fills the arena from the last layer to the first."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    works = []
    for s, e in reversed(list(buckets)):
        view = flat[s:e]
        if flat.is_cuda:
            works.append(dist.all_reduce(view, op=dist.ReduceOp.AVG, group=group, async_op=True))
        else:   # gloo (CPU tests) has no AVG
            works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    if not flat.is_cuda:
        for s, e in buckets:
            flat[s:e].div_(world)


class GradSync:
    """Callable hook object stored on B200ViltModel.grad_sync. With layers_per_chunk > 0 the model issues
    its backward in chunks of that many layers and calls begin / reduce_range / finish so that each
    chunk's gradient span is all-reduced while the layers below it are still computing."""

    def __init__(self, learner, bucket_mb: float = 64.0, group=None, layers_per_chunk: int = 3, sm_reserve=None,
                 defer_to_optimizer: bool = False):
        """sm_reserve: SMs the persistent kernels leave to NCCL's reduction CTAs while an all-reduce is in flight
        (climb_set_sm_reserve); default = $NCCL_MAX_CTAS when that cap is set, else 0 (grids keep every SM).
        defer_to_optimizer: finish() does not make the compute stream wait for the reductions; ArenaAdamW.step() waits span
        by span instead, so the update of the top layers runs while the bottom spans are still on the wire. Only valid when
        NOTHING reads the encoder's gradients between backward() and optimizer.step() (no clipping, no EWC/Fisher pass over
        .grad) -- which is the reference's plain training step (train_vqa.py:168-172)."""
        self.learner = learner
        self.group = group
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.layers_per_chunk = layers_per_chunk
        if sm_reserve is None:
            import os
            sm_reserve = int(os.environ.get("CLIMB_SM_RESERVE", os.environ.get("NCCL_MAX_CTAS", "0")) or 0)
        self.sm_reserve = max(0, int(sm_reserve))
        self.defer_to_optimizer = bool(defer_to_optimizer)
        self._cache = None
        self._hooks = []
        self._works = []
        self._ranges = []
        self._reserved = False
        self._loose = []            # task-head parameters whose gradient autograd has produced in this backward
        self._loose_works = []
        self._callback_queued = False
        vilt = learner.get_encoder().vilt
        vilt.grad_sync = self
        for n, p in learner.named_parameters():
            if not n.startswith(("vilt_encoder.", "viltbert_encoder.")):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._sync_loose))

    # ---- task heads (ordinary autograd tensors outside the arena) -----------------------------------
    def _sync_loose(self, p: torch.Tensor) -> None:
        """post-accumulate-grad hook: remember the parameter. The heads sit between the loss and the encoder, so all of them
        are done when the encoder's backward node starts: begin() sends them as ONE grouped all-reduce that travels under
        the encoder backward. Whatever is left when autograd finishes (encoder frozen / not in the graph) goes out from an
        end-of-backward callback."""
        if p.grad is None:
            return
        self._loose.append(p)
        if not self._callback_queued:
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self._end_of_backward)
                self._callback_queued = True
            except RuntimeError:        # not inside a backward pass (a hook fired by hand)
                self._flush_loose()
                self._wait_loose()

    def _allreduce_views(self, views):
        """Asynchronous mean all-reduce of several gradient views as one NCCL group (one launch instead of one per tensor:
        the host cost of ~50 small collectives per step is what made the adapter configuration launch-bound on 2 GPUs)."""
        if not views:
            return None
        if _FAKE_COMM:
            return _Several([])
        if not views[0].is_cuda:        # gloo (CPU tests) has no AVG and no grouped launch
            return _Several([_SumThenDivide(v, self.group) for v in views])
        if len(views) == 1:
            return dist.all_reduce(views[0], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        with dist._coalescing_manager(self.group, async_ops=True) as cm:
            for v in views:
                dist.all_reduce(v, op=dist.ReduceOp.AVG, group=self.group)
        return cm

    def _flush_loose(self) -> None:
        if not self._loose:
            return
        seen, grads = set(), []
        for p in self._loose:
            if id(p) not in seen and p.grad is not None:
                seen.add(id(p))
                grads.append(p.grad)
        self._loose = []
        w = shard_weight()
        if w != 1.0:
            for g in grads:
                g.mul_(w)
        work = self._allreduce_views(grads)
        if work is not None:
            self._loose_works.append(work)

    def _wait_loose(self) -> None:
        for w in self._loose_works:
            w.wait()
        self._loose_works = []

    def _end_of_backward(self) -> None:
        self._callback_queued = False
        self._flush_loose()
        self._wait_loose()

    # ---- overlapped path -------------------------------------------------------------------------
    def _spans(self, arena):
        items = arena.named_items()
        key = (id(arena.theta), tuple(p.requires_grad for _, p in items))
        if self._cache is None or self._cache[0] != key:
            names = [n for n, _ in items]
            trainable = {n for n, p in items if p.requires_grad}
            spans = trainable_spans(arena.offsets, arena.numels, names, trainable)
            self._cache = (key, bucketize(spans, self.bucket_elems), spans)
        return self._cache[2]

    def layer_chunks(self, n_layers: int):
        """(first, last) layer ranges of the chunked backward, top down: `layers_per_chunk` layers each, except that the
        bottom chunk is cut once more (its upper layers / the lowest layer): what is still on the wire when the backward
        ends is the only part of the exchange nothing overlaps, so the last spans are kept small."""
        c = self.layers_per_chunk
        out, first = [], n_layers - 1
        while first >= 0:
            last = max(0, first - c + 1)
            if last == 0 and first - last >= 1:
                out.append((first, 1))
                out.append((0, 0))
            else:
                out.append((first, last))
            first = last - 1
        return out

    def _reserve(self, on: bool) -> None:
        if self.sm_reserve <= 0 or on == self._reserved:
            return
        from . import _lib
        _lib.climb_set_sm_reserve(self.sm_reserve if on else 0)
        self._reserved = on

    def begin(self, arena) -> None:
        wait_pending(arena)         # a deferred exchange nobody consumed (optimizer step skipped)
        self._works = []
        self._ranges = []
        if dist.get_world_size(self.group) > 1:
            self._flush_loose()     # the task heads' gradients are complete: they travel under the encoder backward

    def reduce_range(self, arena, lo: int, hi: int) -> None:
        """All-reduce (mean) the trainable spans inside [lo, hi) of the gradient arena, asynchronously:
        NCCL's stream waits for the kernels enqueued so far, not for the ones that follow."""
        if dist.get_world_size(self.group) == 1:
            return
        ranges = []
        for s, e in self._spans(arena):
            s2, e2 = max(s, lo), min(e, hi)
            if e2 > s2:
                ranges.extend(bucketize([(s2, e2)], self.bucket_elems))
        if not ranges:
            return
        w = shard_weight()
        if w != 1.0:
            for bs, be in ranges:
                arena.grad[bs:be].mul_(w)
        self._reserve(True)     # from here on NCCL's CTAs sit on some SMs: the persistent grids leave them room
        # large buckets go out one by one (they complete one by one: the deferred optimizer starts on the first while the
        # others travel); a run of small spans (adapters: two per layer) goes out as ONE grouped launch
        small = [(bs, be) for bs, be in ranges if be - bs < self.bucket_elems // 8]
        for bs, be in ranges:
            if (bs, be) not in small or len(small) == 1:
                self._works.append(self._allreduce_views([arena.grad[bs:be]]))
                self._ranges.append((bs, be))
        if len(small) > 1:
            work = self._allreduce_views([arena.grad[bs:be] for bs, be in small])
            for r in small:
                self._works.append(work)
                self._ranges.append(r)

    def finish(self, arena) -> None:
        self._wait_loose()          # issued before the first span: long complete
        if self.defer_to_optimizer:
            # ArenaAdamW.step() consumes these span by span (optim.py); anything else that touches the gradients first
            # calls wait_pending(arena)
            arena._pending_reductions = (list(zip(self._ranges, self._works)), self)
        else:
            for w in self._works:
                w.wait()            # stream-level wait: the optimizer kernels queue behind the reductions
            self._reserve(False)
        self._works, self._ranges = [], []

    # ---- one-shot path (layers_per_chunk = 0) ------------------------------------------------------
    def __call__(self, arena) -> None:
        items = arena.named_items()
        key = (id(arena.theta), tuple(p.requires_grad for _, p in items))
        if self._cache is None or self._cache[0] != key:
            names = [n for n, _ in items]
            trainable = {n for n, p in items if p.requires_grad}
            spans = trainable_spans(arena.offsets, arena.numels, names, trainable)
            self._cache = (key, bucketize(spans, self.bucket_elems), spans)
        w = shard_weight()
        if w != 1.0:
            for s, e in self._cache[1]:
                arena.grad[s:e].mul_(w)
        self._flush_loose()
        allreduce_mean_(arena.grad, self._cache[1], self.group)
        self._wait_loose()

    def detach(self):
        self.learner.get_encoder().vilt.grad_sync = None
        for h in self._hooks:
            h.remove()
        self._hooks = []


import os as _os
_FAKE_COMM = _os.environ.get("CLIMB_FAKE_COMM") == "1"      # measurement only: the chunked backward without its collectives


class _SumThenDivide:
    """gloo stand-in for an async AVG all-reduce (CPU tests): wait() finishes the SUM and divides by the world size."""

    def __init__(self, view, group):
        self.view, self.world = view, dist.get_world_size(group)
        self.work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def wait(self):
        self.work.wait()
        self.view.div_(self.world)


class _Several:
    def __init__(self, works):
        self.works = works

    def wait(self):
        for w in self.works:
            w.wait()


def wait_pending(arena) -> None:
    """Make the current stream wait for a deferred gradient exchange of `arena` (GradSync(defer_to_optimizer=True)) and
    clear it. Called by everything that reads or rewrites the gradient arena outside ArenaAdamW.step()."""
    pend = getattr(arena, "_pending_reductions", None)
    if pend is None:
        return
    for _, w in pend[0]:
        w.wait()
    pend[1]._reserve(False)
    arena._pending_reductions = None


def pending_segments(chunks, ranges):
    """For ArenaAdamW: chunks = [(start, length, group)], ranges = [(lo, hi)] in the order their all-reduces were ISSUED
    (NCCL completes them in that order). Returns (order, bounds): `order` = chunk indices sorted by the index of the LAST
    issued range each chunk overlaps (-1 = none: its gradient is local and final), `bounds[k]` = how many of the sorted
    chunks may be updated once range k has completed (bounds[-1] of the un-ranged prefix is returned as bounds[0] start)."""
    seg = []
    for s, l, _ in chunks:
        k = -1
        for i, (lo, hi) in enumerate(ranges):
            if s < hi and s + l > lo:
                k = i
        seg.append(k)
    order = sorted(range(len(chunks)), key=lambda i: (seg[i], chunks[i][0]))
    n_free = sum(1 for k in seg if k < 0)
    bounds, done = [], n_free
    for i in range(len(ranges)):
        done += sum(1 for k in seg if k == i)
        bounds.append(done)
    return order, n_free, bounds


def attach(learner, bucket_mb: float = 64.0, group=None, layers_per_chunk: int = 3, sm_reserve=None,
           defer_to_optimizer: bool = False) -> GradSync:
    """Make every backward of `learner` end with its gradients averaged over the ranks. Parameters must
    already be identical on all ranks (same seed / same checkpoint), as in any data-parallel run."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    return GradSync(learner, bucket_mb, group, layers_per_chunk, sm_reserve, defer_to_optimizer)


def attach_if_distributed(learner, **kw):
    """What create_*_continual_learner_model calls: under a multi-rank torch.distributed launch (torchrun + the
    unchanged CLiMB driver) attach the gradient synchronisation, make the learner's training forward work on
    this rank's rows (sharded_forward below) and make rank 0 the only writer of checkpoints; a single process gets the
    learner back untouched."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    sync = GradSync(learner, **kw)
    learner._ddp_forward_shard = True
    rank_zero_only_io()
    return sync


# ---- the unchanged harness under torchrun: rows sharded INSIDE the learner's forward ------------------------------------------
# The trainers build plain (unsharded) loaders (vqa_dataset.py:262-268), so under torchrun every rank draws the SAME batch
# (same seed, train_upstream_continual_learning.py:103), and they read the labels from the batch dict themselves --
# NLVR2Trainer.train_step even BEFORE it calls the model (train_nlvr2.py:128-131), so nothing that edits the batch on the way
# into the model can keep inputs and labels consistent for every trainer. Instead the learner's training forward
#   * runs the encoder on this rank's contiguous rows only (NLVR2 image pairs / VCR four-choice tuples are single list entries
#     and stay together),
#   * all-gathers pooled outputs and logits, and returns them for the WHOLE batch -- so the trainer's loss, whatever it does
#     with its labels, is the whole-batch loss on every rank (the logged losses are identical across ranks),
#   * in the backward hands back the gradient of its own rows times the world size: after GradSync's mean over ranks every
#     rank holds exactly the whole-batch gradient, also when the rows do not divide evenly; terms that are identical on all
#     ranks (the EWC penalty) pass through the mean unchanged.
# Evaluation stays replicated (the trainers divide by the full dataset length, train_vqa.py:263).
def rank_rows(n: int, rank: int = None, world: int = None, group=None):
    world = dist.get_world_size(group) if world is None else world
    rank = dist.get_rank(group) if rank is None else rank
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per), per


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, local, n, group):
        world = dist.get_world_size(group)
        lo, hi, per = rank_rows(n, group=group)
        pad = local.new_zeros((per,) + tuple(local.shape[1:]))
        pad[: hi - lo] = local
        if local.is_cuda:
            out = local.new_empty((world * per,) + tuple(local.shape[1:]))
            dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
            parts = [out[r * per: (r + 1) * per] for r in range(world)]
        else:
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad.contiguous(), group=group)
        rows = [rank_rows(n, r, world)[:2] for r in range(world)]
        ctx.lo, ctx.hi, ctx.world = lo, hi, world
        return torch.cat([p[: b - a] for p, (a, b) in zip(parts, rows)], dim=0)

    @staticmethod
    def backward(ctx, g):
        return g[ctx.lo: ctx.hi] * ctx.world, None, None


def sharded_forward(forward_fn, images, texts, group=None):
    """forward_fn(images, texts) -> (pooled [rows, ...], logits [rows, C]) on this rank's rows of the batch; returns both for
    the whole batch (see the block comment above)."""
    n = len(texts)
    lo, hi, _ = rank_rows(n, group=group)
    if hi <= lo:
        raise RuntimeError(f"batch of {n} samples leaves rank {dist.get_rank(group)} of {dist.get_world_size(group)} without rows: "
                           "use a batch size >= the number of ranks")
    pooled, logits = forward_fn(images[lo:hi], texts[lo:hi])
    logits = logits.reshape(hi - lo, -1)            # (the multi-choice head's .squeeze() drops the batch dimension of a 1-row shard)
    return _GatherRows.apply(pooled, n, group), _GatherRows.apply(logits, n, group)


# ---- batch-dict sharding helpers ---------------------------------------------------------------------------------------
# shard_batch_inplace / sharding(converter) cut a collated batch dict to this rank's rows on its way into the model. They
# serve harnesses that read their labels AFTER the conversion (VQATrainer, train_vqa.py:127,156); the registry does NOT
# use them, because NLVR2Trainer reads its labels before (train_nlvr2.py:128-131) -- see sharded_forward below for what
# the unchanged harness gets instead. The learner's train() / eval() announce the mode through set_training_mode().
_training_mode = True
_last_shard = None          # (rows on this rank, rows of the whole batch) of the most recent sharded batch
_SHARD_MARK = "_b200_rank_slice"


def set_training_mode(mode: bool) -> None:
    global _training_mode
    _training_mode = bool(mode)


def shard_batch_inplace(batch: dict, size_key_candidates=("raw_texts", "texts", "images", "labels", "target_scores")) -> dict:
    """Cut every per-sample entry of a collated batch (lists and tensors whose length is the batch size) down to
    this rank's contiguous rows, in place. NLVR2 image pairs / VCR four-choice tuples are single entries of those lists,
    so they stay together. No-op outside a multi-rank run, in evaluation mode, and on a dict that was cut already."""
    global _last_shard
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not _training_mode:
        return batch
    if batch.get(_SHARD_MARK):
        return batch
    n = next((len(batch[k]) for k in size_key_candidates if k in batch and hasattr(batch[k], "__len__")), None)
    if n is None:
        raise KeyError(f"cannot tell the batch size: none of {size_key_candidates} in the batch")
    rank, world = dist.get_rank(), dist.get_world_size()
    per = (n + world - 1) // world
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    if hi <= lo:
        raise RuntimeError(f"batch of {n} samples leaves rank {rank} of {world} without rows: use a batch size >= the "
                           "number of ranks (and drop_last / a multiple of it for exact gradient means)")
    for k, v in list(batch.items()):
        if isinstance(v, str) or not hasattr(v, "__len__") or len(v) != n:
            continue
        batch[k] = v[lo:hi]
    batch[_SHARD_MARK] = True
    _last_shard = (hi - lo, n)
    return batch


def shard_weight() -> float:
    """Factor that turns the mean over ranks of per-rank MEAN-loss gradients into the whole-batch mean when the rows
    do not divide evenly: rows_here * world / rows_total (1.0 for even splits)."""
    if _last_shard is None or not dist.is_initialized():
        return 1.0
    return _last_shard[0] * dist.get_world_size() / _last_shard[1]


def sharding(converter):
    """Wrap a batch2inputs_converter (model_configs[...]['batch2inputs_converter']) so that it first cuts the batch
    dict to this rank's rows."""
    def convert(batch, *a, **k):
        if isinstance(batch, dict):
            shard_batch_inplace(batch)
        return converter(batch, *a, **k)
    convert.__name__ = getattr(converter, "__name__", "convert")
    convert.__wrapped__ = converter
    return convert


_io_patched = False


def rank_zero_only_io() -> None:
    """The driver saves checkpoints from every process (train_upstream_continual_learning.py:235,265-266); with replicated
    parameters the files are identical, so ranks > 0 skip the write instead of racing on the same path. wandb is
    switched off on them as well. json results stay as they are (same content from every rank, a few hundred bytes)."""
    global _io_patched
    if _io_patched or not dist.is_initialized() or dist.get_rank() == 0:
        return
    import os
    os.environ.setdefault("WANDB_MODE", "disabled")
    torch.save = lambda *a, **k: None
    _io_patched = True


def shard_batch(batch: dict, rank: int, world: int, group_size: int = 1) -> dict:
    """Rows of the step's batch owned by `rank`: contiguous slices, keeping NLVR2 image pairs / VCR
    four-choice tuples (group_size rows) together. Tensors and lists are both sliced."""
    def cut(v):
        n = len(v) // group_size
        per = (n + world - 1) // world
        lo, hi = min(n, rank * per) * group_size, min(n, (rank + 1) * per) * group_size
        return v[lo:hi]
    return {k: (cut(v) if hasattr(v, "__len__") and not isinstance(v, str) else v) for k, v in batch.items()}
