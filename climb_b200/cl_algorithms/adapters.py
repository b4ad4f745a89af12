"""AdapterHandler with the reference's surface (src/cl_algorithms/adapters.py:36-65). The reference
resolves `args.adapter_config` through adapter-transformers' AdapterConfig.load; here the two configs
CLiMB ships scripts for are built in, and any dict / AdapterConfig-like object is accepted."""
from __future__ import annotations

import logging

from ..modeling.vilt_model import AdapterSpec

logger = logging.getLogger(__name__)

SUPPORTED_ADAPTER_METHODS = ['vanilla']
ADAPTER_MAP = {'pfeiffer': 'pfeiffer', 'houlsby': 'houlsby'}


class AdapterHandler:
    def __init__(self, adapter_method, args):
        self.args = args
        self.adapter_method = adapter_method
        spec = AdapterSpec.from_config(args.adapter_config)
        spec = AdapterSpec(**vars(spec))
        if getattr(args, "adapter_reduction_factor", 0) > 0:
            spec.reduction_factor = args.adapter_reduction_factor
        self.adapter_config = spec
        logger.info("Adding Adapter layers with configuration: %s", spec)

    def add_adapters_to_model(self, model):
        for task_key in self.args.ordered_cl_tasks:
            model.add_adapter(task_key, config=self.adapter_config)

    def activate_adapter_for_training(self, task_key: str, model):
        model.train_adapter(task_key)
        model.set_active_adapters(task_key)

    def activate_adapter_for_eval(self, task_key: str, model):
        model.set_active_adapters(task_key)
