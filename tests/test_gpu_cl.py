"""GPU tests of the continual-learning pieces on the CUDA path: ArenaAdamW vs torch.optim.AdamW,
EWC (Fisher + penalty + penalty gradient) vs the reference golden, loss kernels vs torch."""
import random
import types

import numpy as np
import pytest
import torch

from oracle import vilt_oracle as vo
from tests.golden_util import ALL_TASKS, TINY, TINY_HW, TINY_T, grad_sample_index, load
from tests.test_gpu_parity import _build, _encodings, _rel

pytestmark = pytest.mark.gpu


def test_loss_kernels_match_torch():
    from climb_b200 import ops
    torch.manual_seed(0)
    logits = (torch.randn(7, 3129, device="cuda") * 3).requires_grad_(True)
    target = torch.zeros(7, 3129, device="cuda")
    target[torch.arange(7), torch.randint(0, 3129, (7,))] = 0.6
    ref = torch.nn.BCEWithLogitsLoss()(logits, target) * 3129
    ref.backward()
    g_ref = logits.grad.clone()
    logits.grad = None
    got = ops.vqa_loss(logits, target)
    (got * 2.0).backward()
    assert abs(got.item() - ref.item()) < 1e-4 * abs(ref.item())
    assert _rel(logits.grad, 2.0 * g_ref) < 1e-5
    lg = torch.randn(9, 4, device="cuda", requires_grad=True)
    tg = torch.randint(0, 4, (9,), device="cuda")
    ref = torch.nn.CrossEntropyLoss()(lg, tg)
    ref.backward()
    g_ref = lg.grad.clone()
    lg.grad = None
    got = ops.cross_entropy_loss(lg, tg)
    got.backward()
    assert abs(got.item() - ref.item()) < 1e-5
    assert _rel(lg.grad, g_ref) < 1e-5


def test_arena_adamw_matches_torch_adamw():
    """Three optimizer steps of the learner's optimizer (arena kernel + loose head tensors) against
    torch.optim.AdamW fed the very same gradients; also checks the bf16 shadow written by the kernel."""
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=11)
    learner = _build(TINY, ALL_TASKS, sd)
    hp = {"lr": 1e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8}
    opt = learner.create_optimizer(hp)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    ref_params = {n: p.detach().clone().requires_grad_(True) for n, p in learner.named_parameters()}
    decay, nodecay = vo.weight_decay_groups(list(ref_params))
    ref_opt = torch.optim.AdamW([{"params": [ref_params[n] for n in decay], "weight_decay": 1e-2},
                                 {"params": [ref_params[n] for n in nodecay], "weight_decay": 0.0}],
                                lr=1e-3, eps=1e-8, betas=(0.9, 0.98))
    ref_sched = torch.optim.lr_scheduler.LambdaLR(ref_opt, lambda s: 1.0 / (1 + s))
    dev = torch.device("cuda")
    for it in range(3):
        batch = vo.synth_batch("snli-ve", 3, TINY, T=TINY_T, image_hw=TINY_HW, seed=50 + it)
        learner.train()
        _, logits = learner.forward_tensors("snli-ve", _encodings("snli-ve", batch, dev))
        torch.nn.CrossEntropyLoss()(logits, batch["target"].to(dev)).backward()
        for n, p in learner.named_parameters():
            ref_params[n].grad = None if p.grad is None else p.grad.detach().clone()
        opt.step(); sched.step(); opt.zero_grad(set_to_none=True)
        ref_opt.step(); ref_sched.step()
    worst = 0.0
    for n, p in learner.named_parameters():
        worst = max(worst, (p.detach() - ref_params[n].detach()).abs().max().item())
    assert worst < 2e-6, worst
    # unused heads (grad None) untouched: no weight decay, as in the reference
    assert torch.equal(dict(learner.named_parameters())["task_layer.vqa.3.weight"].detach().cpu(), sd["task_layer.vqa.3.weight"])
    arena = learner.vilt_encoder.vilt._arena
    assert torch.equal(arena.shadow, arena.theta.bfloat16())


def test_ewc_fisher_and_penalty_vs_reference_golden():
    from climb_b200.cl_algorithms import EWC
    g = load("tiny_ewc_snli-ve")
    seed = int(g["seed"])
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=seed)
    learner = _build(TINY, ALL_TASKS, sd)
    dev = torch.device("cuda")
    sizes = g["batch_sizes"].tolist()
    batches = []
    for i, b in enumerate(sizes):
        bt = vo.synth_batch("snli-ve", b, TINY, T=TINY_T, image_hw=TINY_HW, seed=seed + i, masked=True)
        bt["raw_texts"] = [None] * b
        batches.append(bt)

    class FakeTrainer:
        hparams = {"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8}
        device = dev

        def get_train_dataloader(self):
            class DL(list):
                dataset = [None] * sum(sizes)
            return DL(batches)

        def train_step(self, model, batch, optimizer=None, scheduler=None, ewc=None):
            model.train()
            _, logits = model.forward_tensors("snli-ve", _encodings("snli-ve", batch, dev))
            loss = torch.nn.CrossEntropyLoss()(logits, batch["target"].to(dev))
            loss.backward()
            return loss, logits, None, None

    ewc = EWC(types.SimpleNamespace(ewc_fisher_sample_percentage=1.0, ewc_loss_weight=float(g["ewc_loss_weight"])))
    assert not ewc.do_ewc()
    ewc.save_task_parameters("snli-ve", learner, FakeTrainer(), dev)
    assert ewc.do_ewc()
    arena = learner.vilt_encoder.vilt._arena
    fisher = ewc.fisher_dict["snli-ve"]
    fscale = max(float(g[k]) for k in g.files if k.startswith("fisher_norm/"))
    checked = 0
    for key in g.files:
        if not key.startswith("fisher_norm/"):
            continue
        name = key[len("fisher_norm/vilt."):]
        o, n = arena.offsets[name], arena.numels[name]
        got = fisher[o:o + n].float().cpu()
        ref_norm = float(g[key])
        if ref_norm < 1e-4 * fscale:
            assert got.norm().item() < 1e-2 * fscale, name
            continue
        # Fisher = squares of (cumulative) bf16-path gradients: twice the gradient tolerance
        assert abs(got.norm().item() - ref_norm) <= 0.12 * ref_norm, (name, got.norm().item(), ref_norm)
        checked += 1
    assert checked > 20
    # penalty + gradient at the same perturbed point as the golden run
    gen = torch.Generator().manual_seed(seed + 99)
    with torch.no_grad():
        for n, p in learner.get_encoder().named_parameters():
            p.add_((0.01 * torch.randn(p.shape, generator=gen)).to(dev))
    learner.zero_grad(set_to_none=True)
    # swap in the reference's exact Fisher so that the penalty arithmetic itself is compared tightly
    exact = torch.zeros_like(fisher)
    for key in g.files:
        if key.startswith("fisher/"):
            name = key[len("fisher/vilt."):]
            o, n = arena.offsets[name], arena.numels[name]
            exact[o:o + n] = torch.from_numpy(g[key]).flatten().to(dev)
    have_exact = {k[len("fisher/vilt."):] for k in g.files if k.startswith("fisher/")}
    mask = torch.zeros_like(fisher)
    for name in have_exact:
        o, n = arena.offsets[name], arena.numels[name]
        mask[o:o + n] = 1.0
    ewc.fisher_dict["snli-ve"] = exact * mask
    random.seed(0)
    key_, loss = ewc.compute_ewc_loss(learner)
    loss.backward()
    assert key_ == "snli-ve"
    # reference value restricted to the tensors whose full Fisher is stored in the fixture
    theta = {n: p.detach().float().cpu() for n, p in learner.get_encoder().named_parameters()}
    ref_loss = 0.0
    for name in have_exact:
        f = torch.from_numpy(g["fisher/vilt." + name])
        ref_loss += (f * (theta["vilt." + name] - sd["vilt_encoder.vilt." + name]) ** 2).sum().item()
    ref_loss *= float(g["ewc_loss_weight"])
    assert abs(loss.item() - ref_loss) <= 1e-4 * ref_loss, (loss.item(), ref_loss)
    for name in list(have_exact)[:40]:
        p = dict(learner.get_encoder().named_parameters())["vilt." + name]
        f = torch.from_numpy(g["fisher/vilt." + name])
        ref_g = 2 * float(g["ewc_loss_weight"]) * f * (theta["vilt." + name] - sd["vilt_encoder.vilt." + name])
        if ref_g.norm() < 1e-12:
            continue
        assert _rel(p.grad, ref_g) < 1e-4, name
