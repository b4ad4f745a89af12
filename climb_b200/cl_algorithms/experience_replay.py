"""Experience replay with the reference's hook surface (src/cl_algorithms/experience_replay.py):
host-side index buffers and one extra train_step with a FRESH optimizer per replay step (SURVEY.md
appendix C8). The encoder step it triggers is the CUDA path; `sample_concat_batch` additionally
offers the north-star variant in which replay rows are concatenated to the current batch on the
device (documented deviation: it changes the optimisation semantics, so it is opt-in)."""
from __future__ import annotations

import logging
import random
from typing import Dict

import torch

logger = logging.getLogger(__name__)


class ExperienceReplayMemory:
    def __init__(self):
        self.memory_buffers = {}

    def add_task_memory_buffer(self, args, task_key: str, task_config: Dict, task_trainer, memory_percentage: float,
                               sampling_strategy: str):
        self.memory_buffers[task_key] = TaskMemoryBuffer(args, task_key, task_config, task_trainer, memory_percentage,
                                                         sampling_strategy)

    def do_replay(self) -> bool:
        return True if len(self.memory_buffers) > 0 else False

    def sample_replay_task(self) -> str:
        return random.choice(list(self.memory_buffers.keys()))

    def run_replay_step(self, task_key: str, model) -> torch.Tensor:
        """experience_replay.py:53-67: new AdamW (no moments, base lr, no schedule) + one train_step."""
        task_buffer = self.memory_buffers[task_key]
        task_trainer = task_buffer.task_trainer
        optimizer = model.create_optimizer(task_trainer.hparams)
        replay_batch = task_buffer.sample_replay_batch()
        replay_loss, output, _, _ = task_trainer.train_step(model, replay_batch, optimizer)
        logger.info("%s replay step: loss = %.5f", task_buffer.task_name, float(replay_loss))
        return replay_loss


class TaskMemoryBuffer:
    def __init__(self, args, task_key: str, task_config: Dict, task_trainer, memory_percentage: float,
                 sampling_strategy: str):
        self.task_key = task_key
        self.task_name = task_config.get('task_name', task_key)
        self.task_config = task_config
        self.task_trainer = task_trainer
        self.dataset = task_trainer.get_train_dataloader().dataset
        self.batch_collate_fn = task_trainer.get_collate_fn()
        if task_key == 'nlvr2':
            self.batch_size = int(args.batch_size / 2)
        elif task_key == 'vcr':
            self.batch_size = int(args.batch_size / 4)
        else:
            self.batch_size = args.batch_size
        self.memory_percentage = memory_percentage
        assert self.memory_percentage < 1.0
        self.memory_size = int(memory_percentage * len(self.dataset))
        self.sampling_strategy = sampling_strategy
        assert sampling_strategy in ['random']
        self.memory_idxs = random.sample(list(range(len(self.dataset))), self.memory_size)
        logger.info("Created %s replay memory buffer, with %d samples in the memory", self.task_name, len(self.memory_idxs))

    def __len__(self):
        return len(self.memory_idxs)

    def sample_replay_batch(self) -> Dict:
        sampled_instances = random.sample(self.memory_idxs, self.batch_size)
        return self.batch_collate_fn([self.dataset[i] for i in sampled_instances])


def concat_encodings(current: Dict[str, torch.Tensor], replay: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """North-star ER variant: current || replay rows concatenated ON THE DEVICE before one encoder pass.
    Both must be single-image encodings of the same text length and resolution."""
    out = {}
    for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask"):
        if k in current and k in replay:
            out[k] = torch.cat([current[k], replay[k]], dim=0)
    return out
