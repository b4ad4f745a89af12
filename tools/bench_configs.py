"""Step timings of BASELINE.json configs[2..4] (the parity-test configurations that are NOT bench.py's
headline line) on the CUDA path, one JSON object per config:

  config 3  ViLT-base + Houlsby adapters (rf 16), NLVR2 image pairs, base frozen      (32 pairs = 64 sequences / GPU)
  config 4  ViLT-base Experience Replay: VQA step on current||replay rows concatenated on the device
            (48 current + 16 replay), and the reference-semantics replay step (fresh AdamW, SNLI-VE batch)
  config 5  ViLT-BERT + EWC penalty, VCR 4-choice (16 samples = 64 sequences / GPU), lambda = 100

    python tools/bench_configs.py [--steps 10] [--warmup 3]          (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py   (DDP)

Synthetic tensors, random-init weights, device-resident inputs; fwd + loss (+ EWC) + bwd + AdamW per step,
CUDA events, max over ranks. Dev / evidence tool: results are copied to profiles/."""
from __future__ import annotations

import argparse
import json
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

T_TEXT, IMG = 40, 448
SPECS = {"vqa": dict(num_labels=3129, num_images=1, model_type="classification"),
         "nlvr2": dict(num_labels=2, num_images=2, model_type="classification"),
         "snli-ve": dict(num_labels=3, num_images=1, model_type="classification"),
         "vcr": dict(num_labels=4, num_images=1, model_type="multi-choice", num_choices=4)}
TASKS = list(SPECS)
HP = {"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8}


def text(n, g, dev):
    ids = torch.randint(1000, 30000, (n, T_TEXT), generator=g)
    ids[:, 0], ids[:, -1] = 101, 102
    return {"input_ids": ids.to(dev), "attention_mask": torch.ones(n, T_TEXT, dtype=torch.int64, device=dev),
            "token_type_ids": torch.zeros(n, T_TEXT, dtype=torch.int64, device=dev)}


def pixels(n, g, dev):
    return (torch.rand(n, 3, IMG, IMG, generator=g) * 2 - 1).to(dev)


def timed(fn, steps, warmup, world, dist):
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import torch.distributed as dist
    from climb_b200 import distributed as cdist, ops
    from climb_b200.cl_algorithms.ewc import EWC
    from climb_b200.cl_algorithms.experience_replay import concat_encodings
    from climb_b200.modeling import (B200BertConfig, B200BertModel, B200ViltBertContinualLearner, B200ViltBertEncoderWrapper,
                                     B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel)

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(1234 + rank)
    out = []

    def report(name, seqs, samples, ms, extra=None):
        line = {"config": name, "n_gpus": world, "sequences_per_gpu_step": seqs, "samples_per_gpu_step": samples,
                "ms_per_step": round(ms, 3), "samples_per_s": round(world * samples / ms * 1e3, 1),
                "sequences_per_s": round(world * seqs / ms * 1e3, 1)}
        line.update(extra or {})
        out.append(line)
        if rank == 0:
            print(json.dumps(line), flush=True)

    def vilt_learner():
        torch.manual_seed(42)
        m = B200ViltContinualLearner(TASKS, B200ViltEncoderWrapper(None, B200ViltModel(B200ViltConfig()), dev), 768, SPECS).to(dev)
        m.train()
        return m

    # ---- config 3: adapters, NLVR2 ----
    if a.only in ("", "adapters"):
        m = vilt_learner()
        m.add_adapter("nlvr2", "houlsby")
        m.train_adapter("nlvr2")
        m.set_active_adapters("nlvr2")
        if world > 1:
            cdist.attach(m)
        opt = m.create_optimizer(HP)
        pairs = 32
        enc = dict(text(pairs, g, dev), pixel_values=pixels(2 * pairs, g, dev))
        tgt = torch.randint(0, 2, (pairs,), generator=g).to(dev)
        n_train = sum(p.numel() for p in m.parameters() if p.requires_grad)

        def step():
            _, logits = m.forward_tensors("nlvr2", enc)
            ops.cross_entropy_loss(logits, tgt).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        ms = timed(step, a.steps, a.warmup, world, dist)
        report("config3: ViLT-base + Houlsby adapters (rf 16), NLVR2 pairs, base frozen", 2 * pairs, pairs, ms,
               {"trainable_params": n_train})
        del m, opt

    # ---- config 4: experience replay ----
    if a.only in ("", "er"):
        m = vilt_learner()
        if world > 1:
            cdist.attach(m)
        opt = m.create_optimizer(HP)
        cur = dict(text(48, g, dev), pixel_values=pixels(48, g, dev))
        rep = dict(text(16, g, dev), pixel_values=pixels(16, g, dev))       # rows drawn from the VQA replay buffer
        tgt = torch.zeros(64, 3129, device=dev)
        tgt[torch.arange(64), torch.randint(0, 3129, (64,), generator=g)] = 1.0

        def step_concat():
            enc = concat_encodings(cur, rep)                                  # on-device concat, one encoder pass
            _, logits = m.forward_tensors("vqa", enc)
            ops.vqa_loss(logits, tgt).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        ms = timed(step_concat, a.steps, a.warmup, world, dist)
        report("config4a: ER, current(48)||replay(16) VQA rows concatenated on device, one step", 64, 64, ms)
        sn = dict(text(64, g, dev), pixel_values=pixels(64, g, dev))
        sn_t = torch.randint(0, 3, (64,), generator=g).to(dev)

        def step_replay():                                                    # experience_replay.py:53-67 semantics
            ropt = m.create_optimizer(HP)                                     # fresh AdamW: no moments, base lr
            _, logits = m.forward_tensors("snli-ve", sn)
            ops.cross_entropy_loss(logits, sn_t).backward()
            ropt.step()
            ropt.zero_grad(set_to_none=True)
        ms = timed(step_replay, a.steps, a.warmup, world, dist)
        report("config4b: ER replay step (fresh AdamW) on a 64-row SNLI-VE replay batch", 64, 64, ms)
        del m, opt

    # ---- config 5: ViLT-BERT + EWC, VCR ----
    if a.only in ("", "viltbert"):
        torch.manual_seed(42)
        encw = B200ViltBertEncoderWrapper(None, B200ViltModel(B200ViltConfig()), B200BertModel(B200BertConfig()), dev)
        m = B200ViltBertContinualLearner(TASKS, encw, 768, SPECS).to(dev)
        m.train()
        if world > 1:
            cdist.attach(m)
        opt = m.create_optimizer(HP)
        n = 16
        t4 = text(4 * n, g, dev)
        enc = dict(t4, pixel_values=pixels(n, g, dev))
        tgt = torch.randint(0, 4, (n,), generator=g).to(dev)
        # synthetic Fisher / theta* in the arena layout (F ~ U(0, 1e-3), theta* = theta + N(0, 1e-3))
        arena = m.get_encoder().vilt._arena
        arena.sync(dev)
        ewc = EWC(argparse.Namespace(ewc_fisher_sample_percentage=1.0, ewc_loss_weight=100.0))
        ewc.task_keys.append("vqa")
        ewc.param_dict["vqa"] = arena.theta.detach().clone() + 1e-3 * torch.randn_like(arena.theta)
        ewc.fisher_dict["vqa"] = torch.rand_like(arena.theta) * 1e-3
        ewc.fisher_names["vqa"] = [nm for nm, _ in arena.named_items() if "word_embeddings" not in nm]
        ewc._offsets["vqa"] = dict(arena.offsets)
        random.seed(0)

        def step():
            _, logits = m.forward_tensors("vcr", enc)
            loss = ops.cross_entropy_loss(logits, tgt)
            _, pen = ewc.compute_ewc_loss(m)
            (loss + pen).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        ms = timed(step, a.steps, a.warmup, world, dist)
        report("config5: ViLT-BERT + EWC penalty (lambda 100), VCR 4-choice, frozen BERT in train mode (dropout 0.1 live)",
               4 * n, n, ms)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"bench_configs_{world}gpu.json"), "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
