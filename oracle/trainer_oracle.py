"""CPU restatement of the reference's per-task TRAINER loops -- the code that calls the hot path -- and the
synthetic datasets that drive them. TEST INFRASTRUCTURE: imported by tests/ and by
oracle/make_golden_trainer.py only; nothing under climb_b200/ uses it.

What is restated (each function cites the lines it follows):
  * TaskTrainer.train_step / train / eval of VQATrainer (src/train/visionlanguage_tasks/train_vqa.py:120-282)
    and NLVR2Trainer (train_nlvr2.py:92-239): forward through `model(task_key=..., images=..., texts=...)`,
    BCEWithLogits x num_labels or CrossEntropy, backward, optimizer.step / scheduler.step / zero_grad,
    best-model tracking by copy.deepcopy, VQA score / accuracy evaluation under no_grad;
  * the experience-replay hook inside train() (train_vqa.py:223-226);
  * get_polynomial_decay_schedule_with_warmup(power=1, lr_end=0) (AT/optimization.py:208-263) as a LambdaLR.

Pinned by tests/golden/trainer_*.npz, which oracle/make_golden_trainer.py writes by running the UNMODIFIED
VQATrainer / NLVR2Trainer / ExperienceReplayMemory classes of the reference over the same synthetic datasets
(tests/test_trainer_golden.py replays them on the CPU oracle, tests/test_gpu_zz_trainer.py on the CUDA path).

The datasets are pools of pre-encoded samples: `images` / `texts` handed to the model are opaque handles
(task, index[, image]) and the model's `process_inputs` is replaced by `PoolProcessor`, which gathers the
tensors a ViltProcessor would have produced (the tokenizer vocabulary is not available offline).
"""
from __future__ import annotations

import copy
import random
from typing import Dict, List, Optional

import torch

from . import vilt_oracle as vo

POPULAR_VQA_LABELS = 6          # VQA targets live on the first few labels so that a few steps move the score


# ---------------------------------------------------------------------------------------------------------
# synthetic task datasets
# ---------------------------------------------------------------------------------------------------------
class TaskPool:
    """n pre-encoded samples of one task (what ViltProcessor returns for them) + targets."""

    def __init__(self, task: str, n: int, dims: vo.ViltDims, T: int, hw, seed: int):
        self.task, self.n = task, n
        b = vo.synth_batch(task, n, dims, T=T, image_hw=hw, seed=seed, masked=True)
        self.input_ids, self.attention_mask, self.token_type_ids = b["input_ids"], b["attention_mask"], b["token_type_ids"]
        px = b["pixel_values"]
        self.pixel_values = px if px.dim() == 5 else px[:, None]          # [n, n_img, 3, H, W]
        g = torch.Generator().manual_seed(77_000 + seed)
        spec = vo.TASK_SPECS[task]
        if task == "vqa":
            # soft scores in {0.3, 0.6, 0.9, 1.0} (src/utils/vqa_utils.py:10-20) on 1-2 of the popular labels,
            # label 2 over-represented: a label prior the heads can pick up within a few optimizer steps
            tgt = torch.zeros(n, spec["num_labels"])
            scores = torch.tensor([0.3, 0.6, 0.9, 1.0])
            for i in range(n):
                first = 2 if float(torch.rand((), generator=g)) < 0.6 else int(torch.randint(0, POPULAR_VQA_LABELS, (1,), generator=g))
                tgt[i, first] = scores[int(torch.randint(1, 4, (1,), generator=g))]
                if float(torch.rand((), generator=g)) < 0.5:
                    second = int(torch.randint(0, POPULAR_VQA_LABELS, (1,), generator=g))
                    if second != first:
                        tgt[i, second] = scores[int(torch.randint(0, 2, (1,), generator=g))]
            self.target = tgt
        else:
            p = torch.rand(n, generator=g)
            self.target = torch.where(p < 0.7, torch.ones(n, dtype=torch.long), torch.zeros(n, dtype=torch.long))
            if spec["num_labels"] > 2:
                self.target = torch.randint(0, spec["num_labels"], (n,), generator=g)

    def items(self, lo: int, hi: int) -> List[Dict]:
        """What the datasets' __getitem__ return, reduced to the keys the trainers read."""
        out = []
        for i in range(lo, hi):
            imgs = [(self.task, i, j) for j in range(self.pixel_values.shape[1])]
            # VCR: one text handle per answer choice (forward_multi_choice flattens the list of lists, vilt.py:326-327)
            text = (self.task, i) if self.input_ids.dim() == 2 else [(self.task, i, c) for c in range(self.input_ids.shape[1])]
            item = {"image": imgs[0] if len(imgs) == 1 else imgs, "text": text}
            if self.task == "vqa":
                item["target_scores"] = self.target[i]
            else:
                item["label"] = self.target[i]
            out.append(item)
        return out


def collate(items: List[Dict]) -> Dict:
    """The batch dicts of vqa_dataset.py / nlvr2_dataset.py batch_collate, reduced to what the trainers and
    convert_batch_to_vilt_input_dict (src/modeling/vilt.py:548-553) read."""
    batch = {"images": [it["image"] for it in items], "raw_texts": [it["text"] for it in items]}
    if "target_scores" in items[0]:
        batch["target_scores"] = torch.stack([it["target_scores"] for it in items])
    else:
        batch["labels"] = torch.stack([it["label"] for it in items])
    return batch


class Batches(list):
    """Stand-in for a DataLoader without shuffling: a list of collated batches with the two attributes the
    trainers and TaskMemoryBuffer read (`.dataset`, `.collate_fn`)."""

    def __init__(self, dataset: List[Dict], batch_size: int):
        super().__init__(collate(dataset[i:i + batch_size]) for i in range(0, len(dataset), batch_size))
        self.dataset = dataset
        self.collate_fn = collate


class PoolProcessor:
    """Replaces EncoderWrapper.process_inputs (src/modeling/vilt.py:83-96): handles -> encodings on `device`,
    laid out as ViltProcessor does (images flattened by the caller, one text per sample)."""

    def __init__(self, pools: Dict[str, TaskPool], device):
        self.pools, self.device = pools, device
        self.calls = 0

    def __deepcopy__(self, memo):          # copy.deepcopy(model) must not clone the datasets
        return self

    def __call__(self, images, texts):
        self.calls += 1
        P = self.pools
        pick = lambda field, h: getattr(P[h[0]], field)[h[1]] if len(h) == 2 else getattr(P[h[0]], field)[h[1], h[2]]
        enc = {
            "input_ids": torch.stack([pick("input_ids", h) for h in texts]),
            "attention_mask": torch.stack([pick("attention_mask", h) for h in texts]),
            "token_type_ids": torch.stack([pick("token_type_ids", h) for h in texts]),
            "pixel_values": torch.stack([P[t].pixel_values[i, j] for t, i, j in images]),
        }
        enc["pixel_mask"] = torch.ones(enc["pixel_values"].shape[0], *enc["pixel_values"].shape[-2:], dtype=torch.long)
        return {k: v.to(self.device) for k, v in enc.items()}


# ---------------------------------------------------------------------------------------------------------
# the trainers
# ---------------------------------------------------------------------------------------------------------
def polynomial_decay_lambda(num_warmup_steps: int, num_training_steps: int, lr_init: float, lr_end: float = 0.0,
                            power: float = 1.0):
    """lr_lambda of get_polynomial_decay_schedule_with_warmup (AT/optimization.py:242-261)."""

    def lr_lambda(current_step: int):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        elif current_step > num_training_steps:
            return lr_end / lr_init
        lr_range = lr_init - lr_end
        decay_steps = num_training_steps - num_warmup_steps
        pct_remaining = 1 - (current_step - num_warmup_steps) / decay_steps
        decay = lr_range * pct_remaining ** power + lr_end
        return decay / lr_init

    return lr_lambda


class TrainerOracle:
    """VQATrainer (task 'vqa') / NLVR2Trainer (task 'nlvr2') restated; `record` collects what the golden run
    recorded from the unmodified classes."""

    def __init__(self, task: str, train_dl: Batches, val_dl: Batches, hparams: Dict, num_epochs: int, device,
                 cl_algorithm: str = "sequential_ft", replay_frequency: int = 100):
        assert task in ("vqa", "nlvr2")
        self.task, self.device = task, device
        self.train_dl, self.val_dl = train_dl, val_dl
        self.hparams, self.num_epochs = hparams, num_epochs
        self.cl_algorithm, self.replay_frequency = cl_algorithm, replay_frequency
        self.max_steps = len(train_dl) * num_epochs            # train_vqa.py:96
        self.warmup_ratio = 0.1                                # train_vqa.py:97
        self.loss_criterion = (torch.nn.BCEWithLogitsLoss(reduction="mean") if task == "vqa"
                               else torch.nn.CrossEntropyLoss())                          # train_vqa.py:95, train_nlvr2.py:80
        self.record = {"loss": [], "lr": [], "replay": [], "eval_score": [], "eval_logits": []}
        # model_config['batch2inputs_converter'] (train_vqa.py:62): convert_batch_to_vilt_input_dict, vilt.py:548-553
        self.batch2inputs_converter = lambda batch: {"images": batch["images"], "texts": batch["raw_texts"]}

    # what TaskMemoryBuffer reads (experience_replay.py:86-88)
    def get_train_dataloader(self):
        return self.train_dl

    def get_collate_fn(self):
        return self.train_dl.collate_fn

    def forward_pass(self, model, batch, do_eval=False):
        """train_vqa.py:120-132."""
        inputs = self.batch2inputs_converter(batch)
        if do_eval:
            with torch.no_grad():
                return model(task_key=self.task, **inputs)
        return model(task_key=self.task, **inputs)

    def train_step(self, model, batch, optimizer=None, scheduler=None, ewc=None):
        """train_vqa.py:134-174 / train_nlvr2.py:107-150 (ewc is None on this path)."""
        output = self.forward_pass(model, batch)
        logits = output[1]
        if self.task == "vqa":
            target = batch["target_scores"].to(self.device)
            loss = self.loss_criterion(logits, target) * target.shape[1]
        else:
            loss = self.loss_criterion(logits, batch["labels"].to(self.device))
        loss.backward()
        if optimizer is not None:
            optimizer.step()
            if scheduler is not None:
                scheduler.step()
            optimizer.zero_grad()
        return loss, output, None, None

    def train(self, model, replay_memory=None):
        """train_vqa.py:176-244."""
        model.to(self.device)
        do_replay = self.cl_algorithm == "experience_replay" and replay_memory.do_replay()
        optimizer = model.create_optimizer(self.hparams)
        scheduler = torch.optim.lr_scheduler.LambdaLR(
            optimizer, polynomial_decay_lambda(int(self.max_steps * self.warmup_ratio), self.max_steps, optimizer.defaults["lr"]))
        best_score = 0
        best_model = {"epoch": 0, "model": copy.deepcopy(model)}
        model.zero_grad()
        for epoch in range(self.num_epochs):
            model.train()
            for step, batch in enumerate(self.train_dl):
                self.record["lr"].append(optimizer.param_groups[0]["lr"])
                loss, _, _, _ = self.train_step(model, batch, optimizer, scheduler)
                self.record["loss"].append(float(loss.detach()))
                if do_replay and (step + 1) % self.replay_frequency == 0:
                    replay_task = replay_memory.sample_replay_task()
                    replay_loss = replay_memory.run_replay_step(task_key=replay_task, model=model)
                    self.record["replay"].append((replay_task, float(replay_loss.detach())))
            eval_score = self.eval(model)
            if eval_score > best_score:
                best_score = eval_score
                best_model["epoch"] = epoch
                best_model["model"] = copy.deepcopy(model)
        return best_score, best_model

    def eval(self, model) -> float:
        """train_vqa.py:246-267 (VQA score of the arg-max answer, :99-113) / train_nlvr2.py:213-233 (accuracy)."""
        model.eval()
        eval_score = 0
        all_logits = []
        for batch in self.val_dl:
            logits = self.forward_pass(model, batch, do_eval=True)[1]
            all_logits.append(logits.detach().float().cpu())
            if self.task == "vqa":
                target = batch["target_scores"].to(self.device)
                pred = torch.max(logits, 1)[1]
                one_hots = torch.zeros(*target.size()).to(self.device)
                one_hots.scatter_(1, pred.view(-1, 1), 1)
                eval_score += (one_hots * target).sum(1).sum().item()
            else:
                eval_score += (logits.argmax(-1).cpu() == batch["labels"]).sum().item()
        eval_score = eval_score / len(self.val_dl.dataset) * 100.0
        model.train()
        self.record["eval_score"].append(eval_score)
        self.record["eval_logits"].append(torch.cat(all_logits))
        return eval_score


# ---------------------------------------------------------------------------------------------------------
# a ContinualLearner over the CPU oracle (src/modeling/vilt.py:147-367 surface, as far as the trainers use it)
# ---------------------------------------------------------------------------------------------------------
class OracleLearner:
    def __init__(self, dims: vo.ViltDims, tasks, state_dict: Dict[str, torch.Tensor], process_inputs):
        self.dims, self.tasks = dims, list(tasks)
        self.params = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in state_dict.items()}
        self.process_inputs = process_inputs
        self.training = True

    def named_parameters(self):
        return [(k, v) for k, v in self.params.items() if v.requires_grad]

    def state_dict(self):
        return {k: v.detach() for k, v in self.params.items()}

    def to(self, device):
        return self

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def zero_grad(self):
        for _, p in self.named_parameters():
            p.grad = None

    def create_optimizer(self, hparams):
        """src/modeling/vilt.py:205-215."""
        named = self.named_parameters()
        decay, nodecay = vo.weight_decay_groups([n for n, _ in named])
        d = dict(named)
        groups = [{"params": [d[n] for n in decay], "weight_decay": hparams["weight_decay"]},
                  {"params": [d[n] for n in nodecay], "weight_decay": 0.0}]
        return torch.optim.AdamW(groups, lr=hparams["lr"], eps=hparams["adam_epsilon"], betas=(0.9, 0.98))

    def __call__(self, task_key: str, images, texts):
        spec = vo.TASK_SPECS[task_key]
        n_img = spec["num_images"]
        flat = [h for hs in images for h in hs] if n_img > 1 else images          # vilt.py:280
        n_choices = spec.get("num_choices", 1) if spec["model_type"] == "multi-choice" else 1
        flat_texts = [h for hs in texts for h in hs] if n_choices > 1 else texts   # vilt.py:326-327
        enc = self.process_inputs(flat, flat_texts)
        batch = dict(enc)
        if n_img > 1:
            px = enc["pixel_values"]
            batch["pixel_values"] = px.view(len(texts), n_img, *px.shape[-3:])     # vilt.py:287
        if n_choices > 1:                                                          # vilt.py:329-331
            for k in ("input_ids", "attention_mask", "token_type_ids"):
                batch[k] = enc[k].view(len(texts), n_choices, -1)
        batch.pop("pixel_mask", None)
        return vo.learner_forward(self.params, self.dims, task_key, batch)


# ---------------------------------------------------------------------------------------------------------
# the two scenarios of tests/golden/trainer_*.npz
# ---------------------------------------------------------------------------------------------------------
SCENARIOS = {
    # VQA trained for 3 epochs x 4 steps with experience replay from an NLVR2 memory every 2nd step (learning rates
    # chosen so that the VQA score moves within 12 steps while the trajectory stays well conditioned: with 2e-3 / 1e-3
    # a bf16 forward already changes the replay losses by 2x)
    "trainer_vqa_er": dict(task="vqa", n_train=16, n_val=12, batch_size=4, num_epochs=3, seed=700,
                           hparams={"lr": 1e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8},
                           replay=dict(task="nlvr2", n_train=12, memory_percentage=0.5, replay_frequency=2,
                                       hparams={"lr": 2e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8})),
    # NLVR2 (two images per sample) trained for 3 epochs x 3 steps, sequential fine-tuning
    "trainer_nlvr2": dict(task="nlvr2", n_train=12, n_val=12, batch_size=4, num_epochs=3, seed=701,
                          hparams={"lr": 2e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8}, replay=None),
}


# Scenarios that exist only as trajectories of the UNMODIFIED reference trainers (oracle/make_golden_trainer.py) and are replayed
# by the reference's own trainer code on the CUDA learner (tests/test_gpu_zzz_reference_trainer.py): the two remaining task
# trainers, SNLIVETrainer (train_snli_ve.py) and VCRTrainer (train_vcr.py: four text choices per image). The VCR head's
# Dropout(0.1) (src/modeling/vilt.py:199-202) is set to p = 0 on both sides: it draws from different generators on CPU and GPU.
REFERENCE_SCENARIOS = {
    "trainer_snli_ve": dict(task="snli-ve", n_train=12, n_val=12, batch_size=4, num_epochs=2, seed=702,
                            hparams={"lr": 2e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8}, replay=None),
    # weights scaled as in the single-step VCR fixtures (vilt_oracle.synth_state_dict: without it the four choices of a sample
    # tie at random init and the loss sits at ln 4 whatever the model does)
    # EWC at trainer level: Fisher information of a VQA "previous task" (EWC.save_task_parameters: the loop with un-zeroed
    # cumulative gradients, ewc.py:55-71), then SNLI-VE trained with the penalty added in SNLIVETrainer.train_step
    # (train_snli_ve.py:142-145). The fixture is written with the reference's EWC class on the reference learner; on the GPU the
    # same unmodified trainer runs with climb_b200.cl_algorithms.EWC (device-resident theta*, F) on the CUDA learner.
    "trainer_snli_ve_ewc": dict(task="snli-ve", n_train=12, n_val=8, batch_size=4, num_epochs=2, seed=704,
                                hparams={"lr": 1e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8}, replay=None,
                                ewc=dict(task="vqa", n_train=8, fisher_sample_percentage=1.0, loss_weight=50.0,
                                         hparams={"lr": 1e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8})),
    # Adapters at trainer level: AdapterHandler.add_adapters_to_model / activate_adapter_for_training (adapters.py:52-61, as
    # train_upstream_continual_learning.py:155-160,195-197 calls them) put a Houlsby adapter per task into the model and freeze
    # the ViltModel; NLVR2Trainer.train() then trains the adapter + heads. Reduction factor 4 on the tiny model = bottleneck
    # width 32: on the CUDA learner every site runs through the fused adapter kernel.
    "trainer_nlvr2_adapters": dict(task="nlvr2", n_train=12, n_val=12, batch_size=4, num_epochs=2, seed=705,
                                   hparams={"lr": 5e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8}, replay=None,   # (2e-3: bf16 moves a loss by 2.7 %)
                                   adapters=dict(config="houlsby", reduction_factor=4, tasks=["vqa", "nlvr2"])),
    # (No ViLT-BERT scenario: the reference's ViltBertContinualLearner has no create_optimizer -- viltbert.py defines it on the
    # encoder wrapper only, SURVEY appendix C10 -- so its own trainers cannot run that learner; ViLT-BERT is pinned by the
    # single-step fixtures tiny_viltbert_* instead.)
    "trainer_vcr": dict(task="vcr", n_train=8, n_val=8, batch_size=4, num_epochs=2, seed=703,
                        hparams={"lr": 5e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8}, replay=None,      # (2e-3: bf16 moves the 4th loss by 5 %)
                        scales=dict(layer_scale=6.0, head_scale=20.0)),      # = make_golden.VCR_SCALES
}


def build_data(sc: Dict, dims: vo.ViltDims, T: int, hw):
    """Pools and loaders of a scenario: (pools, train_dl, val_dl, replay_train_dl or None)."""
    task = sc["task"]
    pools = {task: TaskPool(task, sc["n_train"] + sc["n_val"], dims, T, hw, sc["seed"])}
    items = pools[task].items(0, sc["n_train"] + sc["n_val"])
    train_dl = Batches(items[:sc["n_train"]], sc["batch_size"])
    val_dl = Batches(items[sc["n_train"]:], sc["batch_size"])
    replay_dl = None
    prev = sc["replay"] or sc.get("ewc")           # the previous task's loader: replay memory or EWC's Fisher loop
    if prev:
        pools[prev["task"]] = TaskPool(prev["task"], prev["n_train"], dims, T, hw, sc["seed"] + 50)
        replay_dl = Batches(pools[prev["task"]].items(0, prev["n_train"]), sc["batch_size"])
    return pools, train_dl, val_dl, replay_dl


def run_scenario(sc: Dict, model, device, train_dl, val_dl, replay_dl, replay_memory_cls=None):
    """Drive `model` (any ContinualLearner: OracleLearner, B200ViltContinualLearner) through the scenario with the
    restated trainers. `replay_memory_cls` = the ExperienceReplayMemory implementation under test."""
    cl = "experience_replay" if sc["replay"] else "sequential_ft"
    replay_memory = None
    random.seed(sc["seed"])                       # set_seed (train_upstream_continual_learning.py:103) seeds python's RNG too
    if sc["replay"]:
        import types
        r = sc["replay"]
        prev = TrainerOracle(r["task"], replay_dl, replay_dl, r["hparams"], 1, device)
        replay_memory = replay_memory_cls()
        replay_memory.add_task_memory_buffer(args=types.SimpleNamespace(batch_size=sc["batch_size"]), task_key=r["task"],
                                             task_config={"task_name": r["task"]}, task_trainer=prev,
                                             memory_percentage=r["memory_percentage"], sampling_strategy="random")
    trainer = TrainerOracle(sc["task"], train_dl, val_dl, sc["hparams"], sc["num_epochs"], device, cl_algorithm=cl,
                            replay_frequency=sc["replay"]["replay_frequency"] if sc["replay"] else 100)
    best_score, best_model = trainer.train(model, replay_memory=replay_memory)
    rec = dict(trainer.record)
    rec["best_score"], rec["best_epoch"], rec["best_model"] = best_score, best_model["epoch"], best_model["model"]
    rec["trainer"] = trainer
    return rec


def reevaluate_snapshot(rec: Dict) -> float:
    """The deepcopy'd best model (train_vqa.py:210,242) evaluated again must reproduce the evaluation of its epoch:
    largest absolute logit difference against what that epoch's eval() saw. Precision independent -- it checks that
    copy.deepcopy(model) froze the parameters of THAT epoch while training went on."""
    trainer, e = rec["trainer"], rec["best_epoch"]
    n_before = len(trainer.record["eval_score"])
    score = trainer.eval(rec["best_model"])
    logits = trainer.record["eval_logits"].pop()
    trainer.record["eval_score"].pop()
    assert len(trainer.record["eval_score"]) == n_before
    assert score == rec["eval_score"][e], (score, rec["eval_score"][e])
    return (logits - rec["eval_logits"][e]).abs().max().item()
