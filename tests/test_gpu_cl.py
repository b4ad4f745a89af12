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


def test_ewc_backward_leaves_frozen_parameters_and_published_gradients_alone():
    """(a) with a frozen base + adapters no Fisher-tracked parameter trains -- the EWC node's backward must not wipe
    the gradients the encoder already published; (b) with the bottom layers frozen the penalty still COUNTS their terms but
    its gradient reaches the trainable parameters only."""
    from climb_b200.cl_algorithms import EWC
    dev = torch.device("cuda")
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=12)
    learner = _build(TINY, ALL_TASKS, sd)
    arena = learner.vilt_encoder.vilt._arena
    arena.sync(dev)
    ewc = EWC(types.SimpleNamespace(ewc_fisher_sample_percentage=1.0, ewc_loss_weight=10.0))
    names = [n for n, _ in arena.named_items()]
    ewc.task_keys = ["vqa"]
    ewc.param_dict["vqa"] = (arena.theta + 0.01).detach().clone()
    ewc.fisher_dict["vqa"] = torch.full_like(arena.theta, 0.5)
    ewc.fisher_names["vqa"] = names
    ewc._offsets["vqa"] = dict(arena.offsets)
    # (b) bottom layer frozen
    learner.get_encoder().freeze_bottom_k_layers(1)
    learner.zero_grad(set_to_none=True)
    _, loss = ewc.compute_ewc_loss(learner)
    n_real = sum(arena.numels.values())
    assert abs(loss.item() - 10.0 * 0.5 * 1e-4 * n_real) <= 2e-3 * loss.item()         # every tracked parameter counts
    loss.backward()
    frozen = "encoder.layer.0.intermediate.dense.weight"
    live = "encoder.layer.1.intermediate.dense.weight"
    named = dict(arena.named_items())
    assert named[frozen].grad is None and float(arena.grad_view(frozen).abs().max()) == 0.0
    assert torch.allclose(named[live].grad, torch.full_like(named[live], 2 * 10.0 * 0.5 * -0.01), rtol=1e-3)
    # (a) adapters: the encoder publishes adapter gradients, then the EWC node runs with an empty trainable list
    from climb_b200.modeling import AdapterSpec
    learner2 = _build(TINY, ALL_TASKS, sd)
    spec = AdapterSpec.from_config("houlsby")
    spec.reduction_factor = 4
    learner2.add_adapter("vqa", spec)
    learner2.to(dev)
    learner2.train_adapter("vqa")
    arena2 = learner2.vilt_encoder.vilt._arena
    arena2.sync(dev)
    batch = vo.synth_batch("vqa", 2, TINY, T=TINY_T, image_hw=TINY_HW, seed=3)
    ewc2 = EWC(types.SimpleNamespace(ewc_fisher_sample_percentage=1.0, ewc_loss_weight=10.0))
    base_names = [n for n, _ in arena2.named_items() if ".adapters." not in n]
    ewc2.task_keys = ["vqa"]
    ewc2.param_dict["vqa"] = (arena2.theta + 0.01).detach().clone()
    ewc2.fisher_dict["vqa"] = torch.full_like(arena2.theta, 0.5)
    ewc2.fisher_names["vqa"] = base_names
    ewc2._offsets["vqa"] = dict(arena2.offsets)
    learner2.train()
    _, logits = learner2.forward_tensors("vqa", _encodings("vqa", batch, dev))
    task_loss = torch.nn.BCEWithLogitsLoss()(logits, batch["target"].to(dev)) * logits.shape[1]
    task_loss.backward(retain_graph=False)
    ad = next(n for n, _ in arena2.named_items() if ".adapters.vqa.adapter_up.weight" in n)
    g_before = dict(arena2.named_items())[ad].grad.clone()
    assert g_before.abs().max() > 0
    _, pen = ewc2.compute_ewc_loss(learner2)
    pen.backward()
    assert torch.equal(dict(arena2.named_items())[ad].grad, g_before)
    assert all(p.grad is None for n, p in arena2.named_items() if ".adapters." not in n)


def test_arena_adamw_keeps_its_moments_when_the_arena_is_rebuilt():
    """add_adapter / expand_modality_type_embeddings / reallocate_text_image rebuild the flat arena under a live
    optimizer: exp_avg / exp_avg_sq / step are carried over by parameter name."""
    dev = torch.device("cuda")
    sd = vo.synth_state_dict(TINY, ["vqa"], seed=14)
    learner = _build(TINY, ["vqa"], sd)
    opt = learner.create_optimizer({"lr": 1e-3, "weight_decay": 0.0, "adam_epsilon": 1e-8})
    batch = vo.synth_batch("vqa", 2, TINY, T=TINY_T, image_hw=TINY_HW, seed=4)

    seen = {}

    def step():
        _, logits = learner.forward_tensors("vqa", _encodings("vqa", batch, dev))
        (torch.nn.BCEWithLogitsLoss()(logits, batch["target"].to(dev)) * logits.shape[1]).backward()
        seen["g"] = dict(learner.get_encoder().vilt.named_parameters())["encoder.layer.1.output.dense.weight"].grad.flatten().clone()
        opt.step()
        opt.zero_grad(set_to_none=True)

    learner.train()
    step()
    arena = learner.vilt_encoder.vilt._arena
    st = opt._arena_state[id(arena)]
    name = "encoder.layer.1.output.dense.weight"
    o, k = arena.offsets[name], arena.numels[name]
    m_before = st["exp_avg"][o:o + k].clone()
    assert m_before.abs().max() > 0 and st["step"] == 1
    learner.get_encoder().expand_modality_type_embeddings()          # replaces an nn.Embedding: the arena is rebuilt
    old_theta = arena.theta
    arena.sync(dev)
    assert arena.theta is not old_theta
    # the replaced embedding is a new Parameter: the optimizer must be told, as with torch.optim (the trainers build a
    # fresh optimizer per task; here we only check the carried state of the tensors it still owns)
    opt.param_groups[0]["params"] = [p for p in opt.param_groups[0]["params"]
                                     if p.shape != (2, TINY.hidden_size)] + [learner.get_encoder().vilt.embeddings.token_type_embeddings.weight]
    step()
    st2 = opt._arena_state[id(arena)]
    assert st2["step"] == 2 and st2["theta_ptr"] == arena.theta.data_ptr()
    o2 = arena.offsets[name]
    m_after = st2["exp_avg"][o2:o2 + k]
    # exp_avg after the second step = 0.9 * m1 + 0.1 * g2 (a reset would give 0.1 * g2 and step == 1)
    expect = 0.9 * m_before + 0.1 * seen["g"]
    assert ((m_after - expect).norm() / expect.norm()).item() < 1e-5
