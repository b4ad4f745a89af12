"""Adapter bookkeeping behind the surface CLiMB's driver calls (src/cl_algorithms/adapters.py:36-65; call sites
train_upstream_continual_learning.py:155-160, 195-197): one bottleneck adapter per task, added up front, the
current task's adapter trained with the ViltModel frozen, a finished task's adapter re-activated for evaluation.

The reference resolves `args.adapter_config` ('houlsby' / 'pfeiffer' / a json path) through adapter-transformers'
AdapterConfig.load and overrides `reduction_factor` when `args.adapter_reduction_factor > 0` (:41-46). Here the
configuration becomes an `AdapterSpec` -- the fields the ViLT mixins actually execute -- built from the two
presets CLiMB ships scripts for, or from any dict / AdapterConfig-like object; options the CUDA engine does not
implement (compacter / PHM, parallel, invertible, adapter LayerNorm, scaling != 1) raise NotImplementedError
instead of being ignored. The bottleneck itself (x + W_u act(W_d x + b_d) + b_u at one or two sites per layer)
runs inside the engine as tcgen05 GEMMs with swish / relu epilogues (climb_b200/csrc/engine.cu).
"""
from __future__ import annotations

import dataclasses
import logging

from ..modeling.vilt_model import AdapterSpec

logger = logging.getLogger(__name__)

SUPPORTED_ADAPTER_METHODS = ['vanilla']
ADAPTER_MAP = {name: name for name in ('pfeiffer', 'houlsby')}          # presets AdapterSpec.from_config knows


class AdapterHandler:
    def __init__(self, adapter_method, args):
        if adapter_method not in SUPPORTED_ADAPTER_METHODS:
            raise ValueError(f"adapter method {adapter_method!r} is not one of {SUPPORTED_ADAPTER_METHODS}")
        self.args, self.adapter_method = args, adapter_method
        spec = dataclasses.replace(AdapterSpec.from_config(args.adapter_config))      # a private copy of the preset
        override = getattr(args, "adapter_reduction_factor", 0)
        if override and override > 0:
            spec.reduction_factor = override
        self.adapter_config = spec
        logger.info("adapter configuration: %s", spec)

    def add_adapters_to_model(self, model) -> None:
        tasks = list(self.args.ordered_cl_tasks)
        for task_key in tasks:
            model.add_adapter(task_key, config=self.adapter_config)
        logger.info("added adapters for tasks: %s", ", ".join(tasks))

    def activate_adapter_for_training(self, task_key: str, model) -> None:
        """Freeze the ViltModel, unfreeze + activate this task's adapter (task heads stay trainable: appendix C5)."""
        model.train_adapter(task_key)
        model.set_active_adapters(task_key)

    def activate_adapter_for_eval(self, task_key: str, model) -> None:
        model.set_active_adapters(task_key)
