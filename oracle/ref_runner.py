"""Run the UNMODIFIED reference's training step for bench.py's reference arm and its like-for-like GPU baseline.

TEST / BASELINE INFRASTRUCTURE: imported by bench.py's `--impl reference` / `gpu_eager_baseline` legs only, never by
climb_b200/. The reference is imported through oracle/ref_shim.py -- from /root/reference in the build container, from
the archive staged by oracle/stage_ref.py on the GPU box. Everything that computes is the reference's own code:

  * `ViltContinualLearner` over `ViltEncoderWrapper(ViltModel(ViltConfig()))` (src/modeling/vilt.py:30-367, vendored
    transformers 4.17 modeling_vilt.py), random init under seed 42 (train_upstream_continual_learning.py:103);
  * `VQATrainer.train_step` (src/train/visionlanguage_tasks/train_vqa.py:135-174): forward_pass -> BCEWithLogits x 3129 ->
    backward -> optimizer.step -> zero_grad, on a trainer object allocated without its dataset-reading __init__;
  * `ViltContinualLearner.create_optimizer` (vilt.py:205-215): torch AdamW, betas (0.9, 0.98), the reference's grouping.

Only `process_inputs` is replaced (by a function returning the synthetic tensors of BASELINE.json's configs): the PIL /
tokenizer half needs the bert-base-uncased vocabulary, which is not available offline, and is outside the metric.
"""
from __future__ import annotations

import os
import statistics
import time

import torch

T_TEXT, IMG, N_LABELS = 40, 448, 3129


def available() -> bool:
    from . import ref_shim
    return ref_shim.reference_available()


class ReferenceStepper:
    def __init__(self, device: str = "cpu", seed: int = 42):
        from . import ref_shim
        ref_shim.install()
        from transformers import ViltConfig, ViltModel                  # vendored 4.17 fork
        from modeling.vilt import ViltContinualLearner, ViltEncoderWrapper, convert_batch_to_vilt_input_dict
        from configs.task_configs import task_configs
        from train.visionlanguage_tasks.train_vqa import VQATrainer
        self.device = torch.device(device)
        torch.manual_seed(seed)
        enc = ViltEncoderWrapper(ref_shim.StubProcessor(), ViltModel(ViltConfig()), self.device)
        self.learner = ViltContinualLearner(["vqa"], enc, 768, task_configs).to(self.device)
        self.learner.train()
        t = VQATrainer.__new__(VQATrainer)
        torch.nn.Module.__init__(t)
        t.device = self.device
        t.batch2inputs_converter = convert_batch_to_vilt_input_dict
        t.loss_criterion = torch.nn.BCEWithLogitsLoss(reduction="mean")        # train_vqa.py:95
        self.trainer = t
        self.optimizer = self.learner.create_optimizer({"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8})

    def make_batch(self, B: int, seed: int):
        g = torch.Generator().manual_seed(seed)
        ids = torch.randint(1000, 30000, (B, T_TEXT), generator=g)
        ids[:, 0], ids[:, -1] = 101, 102
        enc = {"input_ids": ids, "attention_mask": torch.ones(B, T_TEXT, dtype=torch.int64),
               "token_type_ids": torch.zeros(B, T_TEXT, dtype=torch.int64),
               "pixel_values": torch.rand(B, 3, IMG, IMG, generator=g) * 2 - 1,
               "pixel_mask": torch.ones(B, IMG, IMG, dtype=torch.int64)}
        tgt = torch.zeros(B, N_LABELS)
        for b in range(B):
            k = int(torch.randint(1, 4, (1,), generator=g))
            idx = torch.randperm(N_LABELS, generator=g)[:k]
            tgt[b, idx] = torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (k,), generator=g)]
        enc = {k: v.to(self.device) for k, v in enc.items()}
        return {"images": [None] * B, "raw_texts": [None] * B, "target_scores": tgt, "_enc": enc}

    def step(self, batch, autocast: bool = False):
        """One VQATrainer.train_step (the reference's own method) on `batch`."""
        self.learner.vilt_encoder.process_inputs = lambda images, texts: batch["_enc"]
        # The vendored visual_embed builds its patch-index grid with bare torch.arange / torch.meshgrid (modeling_vilt.py:151-153)
        # and indexes it with indices taken from the pixel mask: on a GPU that mixes a CPU tensor with CUDA indices, which the
        # torch of this image rejects. Running the UNCHANGED code under a default-device context puts those factory calls on
        # the model's device -- the reference itself is not edited.
        import contextlib
        ctx = torch.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()
        with ctx:
            if autocast:
                with torch.autocast(self.device.type, dtype=torch.bfloat16):
                    out = self.trainer.train_step(self.learner, batch, self.optimizer)
            else:
                out = self.trainer.train_step(self.learner, batch, self.optimizer)
        return out[0]

    def time_steps(self, B: int, steps: int, warmup: int, autocast: bool = False):
        """Median seconds per step over `steps` timed steps (after `warmup`), rotating over two batches."""
        batches = [self.make_batch(B, 100 + i) for i in range(2)]
        times = []
        for i in range(warmup + steps):
            if self.device.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            loss = self.step(batches[i % 2], autocast)
            if self.device.type == "cuda":
                torch.cuda.synchronize()
            else:
                float(loss)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return statistics.median(times), times


def cpu_reference(steps: int, warmup: int, batch: int = 4):
    """BASELINE config 1 on the host cores: (samples/s, cores, seconds per step)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    r = ReferenceStepper("cpu")
    sec, _ = r.time_steps(batch, steps, warmup)
    return batch / sec, cores, sec


def gpu_eager(batch: int, steps: int = 3, warmup: int = 2):
    """The same reference module as-is on the GPU (eager): fp32 and bf16 autocast, samples/s each."""
    out = {}
    for name, ac in (("fp32", False), ("bf16_autocast", True)):
        r = ReferenceStepper("cuda")
        sec, _ = r.time_steps(batch, steps, warmup, autocast=ac)
        out[name] = {"samples_per_s": round(batch / sec, 1), "ms_per_step": round(sec * 1e3, 2)}
        del r
        torch.cuda.empty_cache()
    return out
