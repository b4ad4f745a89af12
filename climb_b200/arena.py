"""Flat parameter arena.

All parameters of a module live in ONE contiguous fp32 device buffer (`theta`); each nn.Parameter is
a view into it, so the reference's names / shapes / state_dict keys are untouched (SURVEY.md
appendix B) while the CUDA side sees three flat buffers with identical element offsets:

    theta   fp32   master parameters (what the optimizer updates, what checkpoints store)
    shadow  bf16   tensor-core operand copy, refreshed when theta changes
    grad    fp32   gradient arena; every p.grad is a view into it

That layout is what lets EWC's penalty, Fisher accumulation, AdamW and the DDP all-reduce each be
one streaming kernel / one collective over a single buffer instead of ~230 small tensors
(src/cl_algorithms/ewc.py:82-86 loops over named_parameters; src/modeling/vilt.py:205-215).

The arena re-binds itself whenever torch swaps parameter storage underneath it (`model.to(device)`,
`copy.deepcopy`, `load_state_dict(assign=True)`, a replaced nn.Embedding as in
ViltEncoderWrapper.reallocate_text_image, src/modeling/vilt.py:57-81).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

ALIGN = 64   # elements: 256 B in fp32, 128 B in bf16 (TMA needs 16 B, float4 kernels 16 B)


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class ParamArena:
    def __init__(self, owner: nn.Module, lister: str, with_shadow: bool = True, with_grad: bool = True):
        """owner.<lister>() returns the (name, parameter) list in arena order. The arena keeps a
        reference to its owner (not a closure) so that copy.deepcopy(model) -- which the trainers do
        for best-model tracking, train_vqa.py:210,242 -- yields an independent arena."""
        self.owner = owner
        self.lister = lister
        self.with_shadow = with_shadow
        self.with_grad = with_grad          # False: forward-only parameters (ViLT-BERT's frozen BERT), no gradient arena
        self.theta: Optional[torch.Tensor] = None
        self.shadow: Optional[torch.Tensor] = None
        self.shadow_lo: Optional[torch.Tensor] = None     # bf16(theta - shadow): second half of the bf16x3 split operands
        self._lo_version = None
        self.grad: Optional[torch.Tensor] = None
        self.offsets: Dict[str, int] = {}
        self.numels: Dict[str, int] = {}
        self.size = 0
        self._param_ids: List[int] = []
        self._ptrs: List[int] = []
        self._grad_views: Dict[str, torch.Tensor] = {}
        self._shadow_version = -1
        self._version_sum = 0
        self.shadow_dirty = True

    def __deepcopy__(self, memo):
        import copy
        return ParamArena(copy.deepcopy(self.owner, memo), self.lister, self.with_shadow, self.with_grad)

    def _named_params(self):
        return getattr(self.owner, self.lister)()

    # ---------------------------------------------------------------------------------------
    def _layout_matches(self, items) -> bool:
        if self.theta is None or len(items) != len(self._param_ids):
            return False
        base = self.theta.data_ptr()
        vsum = 0
        for (name, p), pid in zip(items, self._param_ids):
            if id(p) != pid or name not in self.offsets:
                return False
            if p.data_ptr() != base + 4 * self.offsets[name] or p.numel() != self.numels[name]:
                return False
            vsum += p._version
        self._version_sum = vsum      # in-place edits through torch bump the parameters' versions
        return True

    def sync(self, device: Optional[torch.device] = None, allow_cpu: bool = False) -> bool:
        """Make sure every parameter is a view of the arena (rebuild if not). Returns True if rebuilt.
        allow_cpu exists for host-logic tests of the layout only: no kernel accepts a CPU arena."""
        items = list(self._named_params())
        if self._layout_matches(items):
            return False
        if device is None:
            device = items[0][1].device
        if device.type != "cuda" and not allow_cpu:
            raise _lib.ClimbError(
                "climb_b200 parameters must live on a CUDA device before the first forward "
                "(model.to(device)); there is no CPU path")
        offsets, numels, off = {}, {}, 0
        for name, p in items:
            offsets[name] = off
            numels[name] = p.numel()
            off += _round_up(p.numel(), ALIGN)
        size = max(off, ALIGN)
        theta = torch.zeros(size, dtype=torch.float32, device=device)
        old_grads = {}
        with torch.no_grad():
            for name, p in items:
                view = theta[offsets[name]: offsets[name] + p.numel()].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                if p.grad is not None:
                    old_grads[name] = p.grad
                p.data = view
        self.theta, self.offsets, self.numels, self.size = theta, offsets, numels, size
        self.shadow = torch.zeros(size, dtype=torch.bfloat16, device=device) if self.with_shadow else None
        self.shadow_lo, self._lo_version = None, None
        self.grad = torch.zeros(size, dtype=torch.float32, device=device) if self.with_grad else None
        self._grad_views = {}
        for name, p in (items if self.with_grad else []):
            gv = self.grad[offsets[name]: offsets[name] + p.numel()].view(p.shape)
            self._grad_views[name] = gv
            if name in old_grads:           # carry accumulated gradients over (e.g. after .to())
                gv.copy_(old_grads[name].to(device))
                p.grad = gv
        self._param_ids = [id(p) for _, p in items]
        self._version_sum = sum(p._version for _, p in items)
        self.shadow_dirty = True
        return True

    # ---------------------------------------------------------------------------------------
    def refresh_shadow(self) -> None:
        """bf16 copy of theta; skipped when nothing wrote to theta since the last refresh (call
        sync() first). The parameters' torch version counters catch in-place edits made through torch
        (optimizer steps, load_state_dict); our own kernels set shadow_dirty."""
        if self.shadow is None:
            return
        v = self._version_sum
        if self.shadow_dirty or v != self._shadow_version:
            _lib.cast_f32_bf16(self.theta, self.shadow)
            self._shadow_version = v
            self.shadow_dirty = False

    def refresh_shadow_lo(self) -> torch.Tensor:
        """bf16x3 precision mode: (shadow, shadow_lo) = split of theta. Call after refresh_shadow(); recomputed whenever theta
        may have changed since the last split (our own AdamW kernel refreshes only the hi half)."""
        v = (self._version_sum, self.theta.data_ptr(), self._lo_epoch)
        if self.shadow_lo is None or self._lo_version != v:
            if self.shadow_lo is None:
                self.shadow_lo = torch.empty_like(self.shadow)
            _lib.split_f32_bf16x2(self.theta, self.shadow, self.shadow_lo)
            self._lo_version = v
        return self.shadow_lo

    _lo_epoch = 0       # bumped by whoever edits theta behind torch's version counters (the AdamW kernel)

    def touch(self) -> None:
        self._lo_epoch += 1

    def grad_view(self, name: str) -> torch.Tensor:
        return self._grad_views[name]

    def named_items(self):
        return list(self._named_params())

    def prepare_grads(self, trainable: Iterable[Tuple[str, nn.Parameter]]) -> None:
        """Called before a backward writes into the gradient arena (which always ACCUMULATES):
        slices whose p.grad is None (fresh step) or foreign are zeroed; slices that are already
        p.grad keep their contents (gradient accumulation, EWC's un-zeroed Fisher loop ewc.py:55-64)."""
        trainable = list(trainable)
        if not trainable:
            return          # nothing of this arena is written: gradients other nodes already published stay as they are
        if getattr(self, "_pending_reductions", None) is not None:      # a deferred exchange no optimizer step consumed
            from .distributed import wait_pending
            wait_pending(self)
        fresh = [(n, p) for n, p in trainable if not self._is_arena_grad(n, p)]
        if len(fresh) == len(trainable) and not self._any_published():
            self.grad.zero_()           # fresh step, nothing accumulated anywhere in the arena: one memset
        else:
            for n, _ in fresh:
                self._grad_views[n].zero_()

    def _any_published(self) -> bool:
        """Does any parameter of the arena currently hold an arena-backed .grad (accumulated by an earlier backward of
        this step, possibly for parameters outside the caller's `trainable` list)?"""
        return any(self._is_arena_grad(n, p) for n, p in self._named_params() if n in self._grad_views)

    def publish_grads(self, trainable: Iterable[Tuple[str, nn.Parameter]]) -> None:
        """After the backward: hand the arena slices to autograd's .grad fields."""
        for n, p in trainable:
            gv = self._grad_views[n]
            if p.grad is None:
                p.grad = gv
            elif not self._is_arena_grad(n, p):
                p.grad.add_(gv)

    def _is_arena_grad(self, name: str, p: nn.Parameter) -> bool:
        g = p.grad
        return g is not None and g.data_ptr() == self._grad_views[name].data_ptr() and g.is_contiguous()
