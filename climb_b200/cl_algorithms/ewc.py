"""EWC with the same hook surface as src/cl_algorithms/ewc.py (EWC.save_task_parameters /
compute_ewc_loss / do_ewc), re-designed around the flat parameter arena:

  reference (ewc.py)                                   here
  ---------------------------------------------------  ------------------------------------------
  theta*, F kept on the CPU, one tensor per name       one device-resident flat tensor each
  F += grad.pow(2).cpu() per tensor per batch (:61-64)  climb_fisher_accumulate over the grad arena
  penalty: ~4 ATen ops per tensor + 2 x 446 MB H2D     climb_ewc_penalty: one 12 B/param pass; its
  per step (:82-86), autograd backward on top          backward one 20 B/param pass into the grad arena

Quirks reproduced on purpose: gradients are NOT zeroed between the batches of
the Fisher loop (so batch t contributes the square of the cumulative gradient), the sum is divided by
the number of samples, and the penalty uses one uniformly chosen previous task per step via python's
`random` (seeded by set_seed).
"""
from __future__ import annotations

import logging
import random
from typing import Dict, List

import torch

from .. import _lib, ops

logger = logging.getLogger(__name__)


def _encoder_arena(model):
    enc = model.get_encoder()
    vilt = getattr(enc, "vilt", None)
    if vilt is None or not hasattr(vilt, "_arena"):
        raise TypeError("climb_b200 EWC needs a B200 encoder (get_encoder().vilt with a parameter arena)")
    return vilt, vilt._arena


class EWC:
    def __init__(self, args):
        self.fisher_sample_percentage = args.ewc_fisher_sample_percentage
        self.ewc_loss_weight = args.ewc_loss_weight
        self.fisher_dict: Dict[str, torch.Tensor] = {}      # task -> flat F (arena layout)
        self.param_dict: Dict[str, torch.Tensor] = {}       # task -> flat theta*
        self.fisher_names: Dict[str, List[str]] = {}        # task -> encoder parameter names that had a gradient
        self._offsets: Dict[str, Dict[str, int]] = {}
        self.task_keys: List[str] = []

    def save_task_parameters(self, task_key: str, model, task_trainer, device: torch.device):
        """ewc.py:28-73."""
        assert task_key not in self.task_keys
        model.to(device)
        vilt, arena = _encoder_arena(model)
        arena.sync(device)
        self.param_dict[task_key] = arena.theta.detach().clone()
        self.task_keys.append(task_key)

        optimizer = model.create_optimizer(task_trainer.hparams)
        dataloader = task_trainer.get_train_dataloader()
        fisher_sample_size = int(self.fisher_sample_percentage * len(dataloader.dataset))
        self.device = task_trainer.device
        fisher = torch.zeros_like(arena.theta)
        optimizer.zero_grad()
        num_samples_completed = 0
        had_grad = set()
        for step, batch in enumerate(dataloader):
            task_trainer.train_step(model, batch)           # optimizer=None: backward only, grads accumulate
            if arena.theta.data_ptr() != self.param_dict[task_key].data_ptr() and fisher.numel() != arena.size:
                raise _lib.ClimbError("parameter arena changed while accumulating the Fisher information")
            _lib.check(_lib.climb_fisher_accumulate(_lib.ptr(arena.grad), _lib.ptr(fisher), arena.size, _lib.stream()))
            for n, p in arena.named_items():
                if p.grad is not None:
                    had_grad.add(n)
            num_samples_completed += len(batch['raw_texts'])
            if num_samples_completed >= fisher_sample_size:
                break
        _lib.check(_lib.climb_scale_inplace(_lib.ptr(fisher), arena.size, 1.0 / max(1, num_samples_completed), _lib.stream()))
        # names that never received a gradient are not part of the reference's fisher_dict: keep F = 0
        # there AND keep them out of the penalty's gradient bookkeeping (their .grad must stay None)
        keep = torch.zeros_like(fisher)
        for n in had_grad:
            o = arena.offsets[n]
            keep[o:o + arena.numels[n]] = 1.0
        fisher.mul_(keep)
        self.fisher_dict[task_key] = fisher
        self.fisher_names[task_key] = [n for n, _ in arena.named_items() if n in had_grad]
        self._offsets[task_key] = dict(arena.offsets)
        logger.info("Saved encoder parameters for %s task!", task_key)

    def _aligned(self, task_key: str, arena):
        """theta* / F in the arena's CURRENT layout (it changes if adapters were added meanwhile)."""
        saved = self._offsets[task_key]
        if saved == arena.offsets and self.fisher_dict[task_key].numel() == arena.size:
            return self.param_dict[task_key], self.fisher_dict[task_key]
        theta_star = arena.theta.detach().clone()       # new names: theta* = theta, F = 0 -> no penalty
        fisher = torch.zeros_like(arena.theta)
        for n, o_old in saved.items():
            if n in arena.offsets:
                k, o = arena.numels[n], arena.offsets[n]
                theta_star[o:o + k] = self.param_dict[task_key][o_old:o_old + k]
                fisher[o:o + k] = self.fisher_dict[task_key][o_old:o_old + k]
        self.param_dict[task_key], self.fisher_dict[task_key] = theta_star, fisher
        self._offsets[task_key] = dict(arena.offsets)
        return theta_star, fisher

    def compute_ewc_loss(self, model):
        """ewc.py:75-87 -> (sampled task key, ewc_loss_weight * sum F (theta - theta*)^2)."""
        ewc_task_key = random.choice(self.task_keys)
        vilt, arena = _encoder_arena(model)
        arena.sync()
        theta_star, fisher = self._aligned(ewc_task_key, arena)
        named = dict(arena.named_items())
        tracked = [n for n in self.fisher_names[ewc_task_key] if n in named]
        trainable = [(n, named[n]) for n in tracked if named[n].requires_grad]
        fisher_bwd = None
        if len(trainable) != len(tracked):
            # some Fisher-tracked parameters are frozen now (freeze_bottom_k_layers, train_adapter): they still count in
            # the loss, but the backward pass must not write into their gradient slots
            key = (ewc_task_key, fisher.data_ptr(), tuple(n for n, _ in trainable))
            if getattr(self, "_bwd_cache", (None, None))[0] != key:
                fb = torch.zeros_like(fisher)
                for n, _ in trainable:
                    o, k = arena.offsets[n], arena.numels[n]
                    fb[o:o + k] = fisher[o:o + k]
                self._bwd_cache = (key, fb)
            fisher_bwd = self._bwd_cache[1]
        return ewc_task_key, ops.ewc_penalty(arena, theta_star, fisher, self.ewc_loss_weight, trainable, fisher_bwd)

    def do_ewc(self):
        return True if len(self.task_keys) > 0 else False
