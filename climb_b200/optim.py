"""ArenaAdamW: torch.optim.AdamW semantics (what ViltContinualLearner.create_optimizer builds,
src/modeling/vilt.py:205-215) executed as ONE CUDA kernel per flat parameter arena.

It is a real torch.optim.Optimizer -- param_groups, lr schedulers (the trainers wrap it in
get_polynomial_decay_schedule_with_warmup, train_vqa.py:199-205) and zero_grad keep working -- but
step() does not loop over ~230 tensors: parameters that live in a ParamArena are updated by
climb_adamw_step over (theta, grad, exp_avg, exp_avg_sq) arenas with a chunk -> param-group table;
parameters outside any arena (the task heads) go through the same kernel on their own storage.

Reference semantics kept: a parameter whose .grad is None is skipped entirely (no moment update, no
weight decay) -- that is what makes ER's fresh-optimizer replay step (experience_replay.py:53-67)
and the unused task heads behave as in the reference. Bias correction uses one step counter per
arena / per loose parameter, counting the steps in which it actually had a gradient (torch keeps the
counter per parameter; the two agree whenever a tensor's gradient is present from its first step on,
which holds for every CLiMB algorithm because each task builds a fresh optimizer).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib

_CHUNK = 1 << 16
_LOOSE_CHUNK = 1 << 13      # task heads are small tensors launched one by one: smaller chunks = enough CTAs to stream from HBM


def _upload_chunks(chunks, device) -> torch.Tensor:
    arr = (_lib.AdamWChunkC * len(chunks))()
    for i, (s, l, g) in enumerate(chunks):
        arr[i].start, arr[i].length, arr[i].group = s, l, g
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)


class ArenaAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, arenas: Optional[List] = None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) > 8:
            raise _lib.ClimbError("ArenaAdamW supports up to 8 param groups")
        self.arenas = list(arenas or [])
        self._arena_state: Dict[int, Dict] = {}
        self._table_cache = None

    def _deferred_plan(self, aid, a, sig, pending):
        """Chunk table of arena `a` re-sorted by the completion order of the pending all-reduces, and the launch plan
        [(first chunk, end chunk, works to wait for first)]. Cached per (parameter set, range layout)."""
        from .distributed import pending_segments
        ranges = tuple(r for r, _ in pending)
        key = (sig, aid, ranges)
        cache = getattr(self, "_deferred_cache", None)
        if cache is None or cache[0] != key:
            chunks = self._chunk_lists[aid]
            order, n_free, bounds = pending_segments(chunks, ranges)
            table = _upload_chunks([chunks[i] for i in order], a.theta.device)
            # merge the per-range bounds into at most ~6 launches: tiny launches cannot fill the machine
            plan, start, waits, min_chunks = [], 0, [], max(1, len(chunks) // 6)
            if n_free:
                plan.append((0, n_free, []))
                start = n_free
            for k, b in enumerate(bounds):
                waits.append(k)
                if b - start >= min_chunks or k == len(bounds) - 1:
                    plan.append((start, b, list(waits)))
                    start, waits = b, []
            cache = (key, table, plan)
            self._deferred_cache = cache
        works = [w for _, w in pending]
        return cache[1], [(c0, c1, [works[k] for k in ks]) for c0, c1, ks in cache[2]]

    def _arena_of(self, p: torch.Tensor):
        for a in self.arenas:
            if a.theta is None:
                continue
            base = a.theta.data_ptr()
            if base <= p.data_ptr() < base + 4 * a.size:
                return a
        return None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        beta1, beta2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        n_groups = len(self.param_groups)
        lr_arr = (ctypes.c_float * n_groups)(*[float(g["lr"]) for g in self.param_groups])
        wd_arr = (ctypes.c_float * n_groups)(*[float(g["weight_decay"]) for g in self.param_groups])
        in_arena, loose, sig = [], [], []
        for gi, g in enumerate(self.param_groups):
            if tuple(g["betas"]) != (beta1, beta2) or g["eps"] != eps:
                raise _lib.ClimbError("ArenaAdamW needs the same betas / eps in every param group")
            for p in g["params"]:
                grad = p.grad
                if grad is None:
                    continue
                a = self._arena_of(p)
                if a is not None and grad.data_ptr() == a.grad.data_ptr() + (p.data_ptr() - a.theta.data_ptr()):
                    in_arena.append((a, p, gi))
                    sig.append((p.data_ptr(), gi))
                else:
                    loose.append((p, gi))
        stream = _lib.stream()
        sig = tuple(sig)
        if self._table_cache is None or self._table_cache[0] != sig:
            per_arena: Dict[int, List] = {}
            for a, p, gi in in_arena:
                start = (p.data_ptr() - a.theta.data_ptr()) // 4
                n = p.numel()
                lst = per_arena.setdefault(id(a), [a, []])[1]
                for o in range(0, n, _CHUNK):
                    lst.append((start + o, min(_CHUNK, n - o), gi))
            self._table_cache = (sig, {aid: (a, _upload_chunks(ch, a.theta.device), len(ch))
                                       for aid, (a, ch) in per_arena.items()})
            self._chunk_lists = {aid: ch for aid, (a, ch) in per_arena.items()}
            self._deferred_cache = None
        for aid, (a, table, n_chunks) in self._table_cache[1].items():
            pend = getattr(a, "_pending_reductions", None)
            if pend is not None:
                table, launches = self._deferred_plan(aid, a, sig, pend[0])
            else:
                launches = [(0, n_chunks, None)]
            st = self._arena_state.get(aid)
            if st is None or st["theta_ptr"] != a.theta.data_ptr():
                new = dict(exp_avg=torch.zeros_like(a.theta), exp_avg_sq=torch.zeros_like(a.theta),
                           theta_ptr=a.theta.data_ptr(), step=0, layout={n: (a.offsets[n], a.numels[n]) for n in a.offsets})
                if st is not None:
                    # the arena was rebuilt under a live optimizer (add_adapter, reallocate_text_image,
                    # expand_modality_type_embeddings, .to(device)): carry the moments and the step over BY NAME, as
                    # torch.optim keeps its per-parameter state across such edits; tensors that are new or changed shape
                    # start from zero moments
                    for n, (o_new, k_new) in new["layout"].items():
                        old = st["layout"].get(n)
                        if old is not None and old[1] == k_new:
                            new["exp_avg"][o_new:o_new + k_new].copy_(st["exp_avg"][old[0]:old[0] + k_new])
                            new["exp_avg_sq"][o_new:o_new + k_new].copy_(st["exp_avg_sq"][old[0]:old[0] + k_new])
                    new["step"] = st["step"]
                st = new
                self._arena_state[aid] = st
            st["step"] += 1
            # the kernel also refreshes the bf16 shadow of everything it updates; parameters it skipped
            # (grad None) did not change, so the shadow stays coherent if it was coherent before
            coherent = a.shadow is not None and not a.shadow_dirty and a._shadow_version == a._version_sum
            # one launch per arena -- or, behind a deferred gradient exchange (GradSync(defer_to_optimizer=True)), one launch
            # per group of chunks whose all-reduce has completed: the update of the top layers runs while the bottom spans
            # are still on the wire (the chunk table is sorted by completion order; a launch takes a slice of it)
            for c0, c1, works in launches:
                for w in works or ():
                    w.wait()                    # stream-level wait for the reductions this slice depends on
                if c1 > c0:
                    _lib.check(_lib.climb_adamw_step(_lib.ptr(a.theta), _lib.ptr(a.grad), _lib.ptr(st["exp_avg"]),
                                                     _lib.ptr(st["exp_avg_sq"]), _lib.ptr(a.shadow) if coherent else None,
                                                     _lib.ptr(table) + c0 * ctypes.sizeof(_lib.AdamWChunkC), c1 - c0, lr_arr, wd_arr,
                                                     n_groups, beta1, beta2, eps, st["step"], stream))
            if pend is not None:
                pend[1]._reserve(False)         # NCCL's CTAs are gone: the persistent kernels take every SM again
                a._pending_reductions = None
            if not coherent:
                a.shadow_dirty = True
            a.touch()           # theta changed behind torch's version counters: the bf16x3 lo half must be re-split
        for p, gi in loose:
            if not (p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32 and p.is_cuda):
                raise _lib.ClimbError("ArenaAdamW handles contiguous fp32 CUDA parameters")
            st = self.state[p]
            if not st or st.get("group") != gi:
                n = p.numel()
                chunks = [(o, min(_LOOSE_CHUNK, n - o), gi) for o in range(0, n, _LOOSE_CHUNK)]
                st.setdefault("exp_avg", torch.zeros_like(p, memory_format=torch.contiguous_format))
                st.setdefault("exp_avg_sq", torch.zeros_like(p, memory_format=torch.contiguous_format))
                st.setdefault("step", 0)
                st["table"], st["n_chunks"], st["group"] = _upload_chunks(chunks, p.device), len(chunks), gi
            st["step"] += 1
            _lib.check(_lib.climb_adamw_step(_lib.ptr(p), _lib.ptr(p.grad), _lib.ptr(st["exp_avg"]),
                                             _lib.ptr(st["exp_avg_sq"]), None, _lib.ptr(st["table"]), st["n_chunks"],
                                             lr_arr, wd_arr, n_groups, beta1, beta2, eps, st["step"], stream))
        return loss
