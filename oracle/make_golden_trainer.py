"""Generate tests/golden/trainer_*.npz by running the UNMODIFIED reference trainers -- VQATrainer.train /
.train_step / .eval (src/train/visionlanguage_tasks/train_vqa.py), NLVR2Trainer (train_nlvr2.py),
ExperienceReplayMemory / TaskMemoryBuffer (src/cl_algorithms/experience_replay.py) and
ViltContinualLearner.create_optimizer + get_polynomial_decay_schedule_with_warmup -- over the synthetic
datasets of oracle/trainer_oracle.py on CPU.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_trainer

TEST INFRASTRUCTURE (build container only: needs /root/reference). The trainer objects are allocated without
running their __init__ (which opens the COCO / VQA / NLVR2 files) and given the attributes __init__ would
have set; every method that then runs -- train, train_step, forward_pass, eval, compute_score_with_logits,
run_replay_step, sample_replay_batch -- is the reference's own code. The model is the reference's
ViltContinualLearner with process_inputs replaced by the pool lookup (no tokenizer vocabulary offline).
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import trainer_oracle as to  # noqa: E402
from oracle.make_golden import (ALL_TASKS, GOLDEN_DIR, TINY, TINY_BERT, TINY_HW, TINY_T, build_reference_learner,  # noqa: E402
                                build_reference_viltbert, grad_sample_index)
from oracle.vilt_oracle import synth_state_dict, synth_viltbert_state_dict  # noqa: E402

PARAM_FULL_MAX = 4096
ALL_SCENARIOS = dict(to.SCENARIOS, **to.REFERENCE_SCENARIOS)


def make_reference_trainer(task, train_dl, val_dl, hparams, num_epochs, cl_algorithm, replay_frequency, record, device="cpu",
                           converter=None):
    from modeling.vilt import convert_batch_to_vilt_input_dict
    if converter is not None:       # what TaskTrainer.__init__ reads from model_configs[args.encoder_name]['batch2inputs_converter']
        convert_batch_to_vilt_input_dict = converter
    if task == "vqa":
        from train.visionlanguage_tasks.train_vqa import VQATrainer as cls
        loaders = dict(vqa_train_dataloader=train_dl, vqa_val_dataloader=val_dl)
        crit = torch.nn.BCEWithLogitsLoss(reduction="mean")
    elif task == "snli-ve":
        from train.visionlanguage_tasks.train_snli_ve import SNLIVETrainer as cls
        loaders = dict(snli_ve_train_dataloader=train_dl, snli_ve_dev_dataloader=val_dl)
        crit = torch.nn.CrossEntropyLoss()
    elif task == "vcr":
        from train.visionlanguage_tasks.train_vcr import VCRTrainer as cls
        loaders = dict(vcr_train_dataloader=train_dl, vcr_val_dataloader=val_dl)
        crit = torch.nn.CrossEntropyLoss()
    else:
        from train.visionlanguage_tasks.train_nlvr2 import NLVR2Trainer as cls
        loaders = dict(nlvr_train_dataloader=train_dl, nlvr_val_dataloader=val_dl)
        crit = torch.nn.CrossEntropyLoss()
    t = cls.__new__(cls)
    torch.nn.Module.__init__(t)                      # TaskTrainer is an nn.Module (task_trainer.py:5-9)
    t.args = types.SimpleNamespace(cl_algorithm=cl_algorithm, replay_frequency=replay_frequency)
    t.device = torch.device(device)
    t.batch2inputs_converter = convert_batch_to_vilt_input_dict
    for k, v in loaders.items():
        setattr(t, k, v)
    t.num_epochs, t.hparams, t.loss_criterion = num_epochs, hparams, crit
    t.max_steps, t.warmup_ratio = len(train_dl) * num_epochs, 0.1
    # recorders around the unmodified methods (instance attributes shadow the class's functions)
    ref_train_step, ref_eval, ref_forward = cls.train_step, cls.eval, cls.forward_pass

    def train_step(model, batch, optimizer=None, scheduler=None, ewc=None):
        if scheduler is not None:
            record["lr"].append(optimizer.param_groups[0]["lr"])
        out = ref_train_step(t, model, batch, optimizer, scheduler, ewc)
        if scheduler is not None:
            record["loss"].append(float(out[0]))
        return out

    def forward_pass(model, batch, do_eval=False):
        out = ref_forward(t, model, batch, do_eval)
        if do_eval:
            record["_eval_tmp"].append(out[1].detach().clone().cpu())
        return out

    def eval_(model):
        record["_eval_tmp"] = []
        score = ref_eval(t, model)
        record["eval_score"].append(score)
        record["eval_logits"].append(torch.cat(record.pop("_eval_tmp")))
        return score

    t.train_step, t.eval, t.forward_pass = train_step, eval_, forward_pass
    return t


def scenario_state_dict(sc):
    """Seeded weights of a scenario: (base state dict, full state dict incl. the adapters' weights when the scenario has any)."""
    kw = dict(sc.get("scales", {}))
    if sc.get("encoder") == "viltbert":
        base = synth_viltbert_state_dict(TINY, TINY_BERT, ALL_TASKS, seed=sc["seed"], **kw)
        return base, base
    base = synth_state_dict(TINY, ALL_TASKS, seed=sc["seed"], **kw)
    if not sc.get("adapters"):
        return base, base
    a = sc["adapters"]
    r = TINY.hidden_size // a["reduction_factor"]
    sites = ("mh", "output") if a["config"] == "houlsby" else ("output",)
    full = synth_state_dict(TINY, ALL_TASKS, seed=sc["seed"], adapters={t: r for t in a["tasks"]}, adapter_sites=sites, **kw)
    return {k: v for k, v in full.items() if ".adapters." not in k}, full


def prepare_adapters(sc, learner, handler_cls):
    """What the driver does before a task with --cl_algorithm adapter (train_upstream_continual_learning.py:155-160,195-197):
    handler.add_adapters_to_model(model), then activate_adapter_for_training(task, model) -- with the seeded adapter weights
    loaded in between so that the reference model and the CUDA model start from the same bottlenecks."""
    a = sc["adapters"]
    handler = handler_cls("vanilla", types.SimpleNamespace(adapter_config=a["config"], adapter_reduction_factor=a["reduction_factor"],
                                                           ordered_cl_tasks=a["tasks"]))
    handler.add_adapters_to_model(learner)
    _, full = scenario_state_dict(sc)
    missing, unexpected = learner.load_state_dict(full, strict=False)
    assert not unexpected, unexpected
    handler.activate_adapter_for_training(sc["task"], learner)
    return handler


def run_reference_scenario(tag, learner, device="cpu", ewc_cls=None, converter=None):
    """Drive `learner` -- the reference's own ViltContinualLearner when the golden trajectories are written, the CUDA learner in
    tests/test_gpu_zzz_reference_trainer.py -- through scenario `tag` with the UNMODIFIED reference trainers and the UNMODIFIED
    ExperienceReplayMemory. Returns (record in oracle.trainer_oracle.run_scenario's format, extras)."""
    from cl_algorithms.experience_replay import ExperienceReplayMemory
    import cl_algorithms.experience_replay as er_mod
    import train.visionlanguage_tasks.train_vqa as tv
    import train.visionlanguage_tasks.train_nlvr2 as tn
    import train.visionlanguage_tasks.train_snli_ve as ts
    import train.visionlanguage_tasks.train_vcr as tc
    tv.tqdm = tn.tqdm = ts.tqdm = tc.tqdm = lambda it, **k: it
    sc = ALL_SCENARIOS[tag]
    dims = TINY
    if "vcr" in getattr(learner, "task_layer", {}):
        learner.task_layer["vcr"][0].p = 0.0         # the head's Dropout(0.1): see trainer_oracle.REFERENCE_SCENARIOS
    pools, train_dl, val_dl, replay_dl = to.build_data(sc, dims, TINY_T, TINY_HW)
    proc = to.PoolProcessor(pools, torch.device(device))
    if hasattr(learner, "processor"):            # oracle.trainer_oracle.OracleLearner (CPU debugging of the harness logic)
        learner.processor = proc
    else:
        learner.get_encoder().process_inputs = proc
    record = {"loss": [], "lr": [], "replay": [], "eval_score": [], "eval_logits": [], "ewc": []}
    cl = "experience_replay" if sc["replay"] else ("ewc" if sc.get("ewc") else "sequential_ft")
    replay_memory = None
    ewc = None
    memory_idxs, sampled = [], []
    random.seed(sc["seed"])
    torch.manual_seed(sc["seed"])                    # visual_embed's multinomial permutation
    ref_sample = er_mod.TaskMemoryBuffer.sample_replay_batch
    if sc["replay"]:
        r = sc["replay"]
        prev_rec = {"loss": [], "lr": [], "eval_score": [], "eval_logits": []}
        prev = make_reference_trainer(r["task"], replay_dl, replay_dl, r["hparams"], 1, cl, 0, prev_rec, device, converter)
        replay_memory = ExperienceReplayMemory()
        replay_memory.add_task_memory_buffer(args=types.SimpleNamespace(batch_size=sc["batch_size"]), task_key=r["task"],
                                             task_config={"task_name": r["task"]}, task_trainer=prev,
                                             memory_percentage=r["memory_percentage"], sampling_strategy="random")
        memory_idxs = list(replay_memory.memory_buffers[r["task"]].memory_idxs)
        ref_replay = ExperienceReplayMemory.run_replay_step

        def sample_replay_batch(self):
            b = ref_sample(self)
            sampled.append([h[1] for h in b["raw_texts"]])
            return b

        er_mod.TaskMemoryBuffer.sample_replay_batch = sample_replay_batch

        def run_replay_step(task_key, model):
            loss = ref_replay(replay_memory, task_key=task_key, model=model)
            record["replay"].append((task_key, float(loss)))
            return loss

        replay_memory.run_replay_step = run_replay_step
    if sc.get("ewc"):
        e = sc["ewc"]
        if ewc_cls is None:
            from cl_algorithms.ewc import EWC as ewc_cls          # the reference's own
            import cl_algorithms.ewc as ewc_mod
            ewc_mod.tqdm = lambda it, **k: it
        prev_rec = {"loss": [], "lr": [], "eval_score": [], "eval_logits": []}
        prev = make_reference_trainer(e["task"], replay_dl, replay_dl, e["hparams"], 1, cl, 0, prev_rec, device, converter)
        ewc = ewc_cls(types.SimpleNamespace(ewc_fisher_sample_percentage=e["fisher_sample_percentage"], ewc_loss_weight=e["loss_weight"]))
        ewc.save_task_parameters(task_key=e["task"], model=learner, task_trainer=prev, device=torch.device(device))
        ref_penalty = ewc.compute_ewc_loss

        def compute_ewc_loss(model):
            task, loss = ref_penalty(model)
            record["ewc"].append((task, float(loss)))
            return task, loss

        ewc.compute_ewc_loss = compute_ewc_loss
    trainer = make_reference_trainer(sc["task"], train_dl, val_dl, sc["hparams"], sc["num_epochs"], cl,
                                     sc["replay"]["replay_frequency"] if sc["replay"] else 100, record, device, converter)
    try:
        best_score, best_model = trainer.train(learner, replay_memory=replay_memory, ewc=ewc)
    finally:
        er_mod.TaskMemoryBuffer.sample_replay_batch = ref_sample
    rec = dict(record)
    rec["best_score"], rec["best_epoch"], rec["best_model"], rec["trainer"] = best_score, best_model["epoch"], best_model["model"], trainer
    return rec, dict(proc=proc, memory_idxs=memory_idxs, sampled=sampled, sc=sc)


def run(tag):
    sc = ALL_SCENARIOS[tag]
    sd, _ = scenario_state_dict(sc)
    learner = build_reference_viltbert(TINY, TINY_BERT, ALL_TASKS, sd) if sc.get("encoder") == "viltbert" else \
        build_reference_learner(TINY, ALL_TASKS, sd)
    if sc.get("adapters"):
        from cl_algorithms.adapters import AdapterHandler         # the reference's own
        prepare_adapters(sc, learner, AdapterHandler)
    record, extra = run_reference_scenario(tag, learner)
    proc, memory_idxs, sampled = extra["proc"], extra["memory_idxs"], extra["sampled"]
    best_score, best_model = record["best_score"], {"epoch": record["best_epoch"], "model": record["best_model"]}
    out = {"loss": np.array(record["loss"], np.float64), "lr": np.array(record["lr"], np.float64),
           "eval_score": np.array(record["eval_score"], np.float64), "best_score": np.float64(best_score),
           "best_epoch": np.int64(best_model["epoch"]), "seed": np.int64(sc["seed"]),
           "process_inputs_calls": np.int64(proc.calls)}
    if sc.get("adapters"):
        out["trainable"] = np.array(sorted(n for n, p in learner.named_parameters() if p.requires_grad))
    for e, lg in enumerate(record["eval_logits"]):
        out[f"eval_logits/{e}"] = lg.numpy()
    if sc.get("ewc"):
        out["ewc_loss"] = np.array([l for _, l in record["ewc"]], np.float64)
        out["ewc_task"] = np.array([t for t, _ in record["ewc"]])
    if sc["replay"]:
        out["replay_loss"] = np.array([l for _, l in record["replay"]], np.float64)
        out["replay_task"] = np.array([t for t, _ in record["replay"]])
        out["memory_idxs"] = np.array(memory_idxs, np.int64)
        out["replay_samples"] = np.array(sampled, np.int64)
    for which, m in (("final", learner), ("best", best_model["model"])):
        for n, p in m.named_parameters():
            v = p.detach()
            out[f"{which}_norm/{n}"] = np.float64(v.double().norm().item())
            if which == "final":
                flat = v.flatten().numpy()
                out[f"final_sample/{n}"] = flat.copy() if flat.size <= PARAM_FULL_MAX else flat[grad_sample_index(flat.size)].copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{tag}.npz"), **out)
    print(f"{tag}: loss {record['loss'][0]:.4f} -> {record['loss'][-1]:.4f}  eval {record['eval_score']}  best epoch "
          f"{best_model['epoch']}  replay {record['replay']}")


def main():
    ref_shim.install()
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for tag in ALL_SCENARIOS:
        if not only or tag in only:
            run(tag)


if __name__ == "__main__":
    main()
