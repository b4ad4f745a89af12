"""Experience replay behind the hook surface CLiMB's driver and trainers call
(src/cl_algorithms/experience_replay.py; call sites train_upstream_continual_learning.py:205-214 and
train_vqa.py:186-188, 223-226):

    memory = ExperienceReplayMemory()
    memory.add_task_memory_buffer(args, task_key, task_config, task_trainer, memory_percentage, sampling_strategy)
    if memory.do_replay(): task = memory.sample_replay_task(); memory.run_replay_step(task_key=task, model=model)

The bookkeeping is host-side index arithmetic; the encoder step a replay triggers is the CUDA path. Two things
are kept exactly, because the reference's results depend on them:

  * the python `random` stream is consumed in the reference's order -- one random.sample over the training
    indices when a buffer is created (:104-106), one random.choice per replay task (:50), one random.sample
    over the memory per replay batch (:120) -- so a seeded run replays the SAME samples as the reference
    (pinned by tests/golden/trainer_vqa_er.npz, recorded from the unmodified classes);
  * a replay step is a separate optimizer step with a FRESH AdamW from model.create_optimizer(hparams): no
    moments, base learning rate, no schedule (:61-63).

`concat_encodings` is the north-star variant (BASELINE.json): replay rows concatenated to the current batch on
the device and pushed through ONE encoder pass. It changes the optimisation semantics (one step instead of
two), so nothing here uses it implicitly; tools/bench_configs.py times it for BASELINE config 4.
"""
from __future__ import annotations

import logging
import random
from typing import Dict, List

import torch

logger = logging.getLogger(__name__)

# sequences per sample: a replay batch holds batch_size / this many samples, so that every replay step
# pushes the same number of sequences through the encoder (experience_replay.py:93-98)
_SEQUENCES_PER_SAMPLE = {"nlvr2": 2, "vcr": 4}
_SAMPLING_STRATEGIES = ("random",)


class TaskMemoryBuffer:
    """Indices of the training samples of one finished task that may be replayed later."""

    def __init__(self, args, task_key: str, task_config: Dict, task_trainer, memory_percentage: float,
                 sampling_strategy: str):
        loader = task_trainer.get_train_dataloader()
        self.task_key, self.task_config, self.task_trainer = task_key, task_config, task_trainer
        self.task_name = task_config.get("task_name", task_key)
        self.dataset, self.batch_collate_fn = loader.dataset, task_trainer.get_collate_fn()
        self.batch_size = int(args.batch_size / _SEQUENCES_PER_SAMPLE.get(task_key, 1))
        assert memory_percentage < 1.0, "memory_percentage is a fraction of the training set"
        assert sampling_strategy in _SAMPLING_STRATEGIES, f"sampling strategy {sampling_strategy!r} is not implemented"
        self.memory_percentage, self.sampling_strategy = memory_percentage, sampling_strategy
        n_train = len(self.dataset)
        self.memory_size = int(memory_percentage * n_train)
        self.memory_idxs: List[int] = random.sample(list(range(n_train)), self.memory_size)
        logger.info("%s: replay memory of %d / %d training samples", self.task_name, self.memory_size, n_train)

    def __len__(self) -> int:
        return len(self.memory_idxs)

    def sample_replay_batch(self) -> Dict:
        chosen = random.sample(self.memory_idxs, self.batch_size)
        return self.batch_collate_fn([self.dataset[i] for i in chosen])


class ExperienceReplayMemory:
    """task_key -> TaskMemoryBuffer, in the order the tasks were learned."""

    def __init__(self):
        self.memory_buffers: Dict[str, TaskMemoryBuffer] = {}

    def add_task_memory_buffer(self, args, task_key: str, task_config: Dict, task_trainer, memory_percentage: float,
                               sampling_strategy: str) -> None:
        self.memory_buffers[task_key] = TaskMemoryBuffer(args, task_key, task_config, task_trainer,
                                                         memory_percentage, sampling_strategy)

    def do_replay(self) -> bool:
        return len(self.memory_buffers) > 0

    def sample_replay_task(self) -> str:
        return random.choice(list(self.memory_buffers))

    def run_replay_step(self, task_key: str, model) -> torch.Tensor:
        buffer = self.memory_buffers[task_key]
        trainer = buffer.task_trainer
        fresh_optimizer = model.create_optimizer(trainer.hparams)
        loss = trainer.train_step(model, buffer.sample_replay_batch(), fresh_optimizer)[0]
        logger.info("%s replay step: loss = %.5f", buffer.task_name, loss.item())
        return loss


def concat_encodings(current: Dict[str, torch.Tensor], replay: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Current-task rows || replay rows on the device, for one encoder pass over both (north-star ER variant).
    Both must be single-image encodings with the same text length and image size."""
    keys = [k for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")
            if k in current and k in replay]
    for k in keys:
        if current[k].shape[1:] != replay[k].shape[1:]:
            raise ValueError(f"{k}: current rows {tuple(current[k].shape[1:])} and replay rows "
                             f"{tuple(replay[k].shape[1:])} differ; pad both to a common text length / image size first")
    return {k: torch.cat([current[k], replay[k]], dim=0) for k in keys}
