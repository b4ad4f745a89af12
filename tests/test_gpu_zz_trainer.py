"""GPU: the reference's TRAINER loops drive the CUDA path end to end -- B200ViltContinualLearner called as
`model(task_key=..., images=..., texts=...)`, ArenaAdamW from `create_optimizer`, the polynomial-decay
schedule, copy.deepcopy best-model snapshots, eval() under no_grad (the save-nothing forward), and this repo's
ExperienceReplayMemory firing replay steps with fresh optimizers -- and the whole trajectory is compared with
what the UNMODIFIED VQATrainer / NLVR2Trainer / ExperienceReplayMemory recorded on the reference model
(tests/golden/trainer_*.npz, oracle/make_golden_trainer.py). The loops themselves are the restatement in
oracle/trainer_oracle.py, pinned to the same fixtures on CPU by tests/test_trainer_golden.py.

(File name: pytest runs files alphabetically; these multi-step scenarios come after the per-kernel and
single-step parity files.)

Stated tolerances (bf16 tensor-core operands; 12 / 9 optimizer steps + replay steps on a tiny model at lr 1e-3 /
2e-3, where Adam turns every rounding difference in a small gradient into a +-lr step):
  training and replay losses   relative error <= 2e-2 / 4e-2
  evaluation logits            relative Frobenius error <= 5e-2; arg-max decisions, VQA score / accuracy and the
                               best epoch EQUAL to the reference's except where the reference's own top-2 margin
                               is below twice the largest logit difference (near ties)
  final parameters             per tensor ||theta - theta_ref|| <= 1.5 ||theta_ref - theta_init||, median over
                               tensors <= 0.3 (a CPU bf16-autocast run of the reference arithmetic sits at
                               0.09 median / 0.7 worst on the VQA + replay scenario, 0.02 / 0.11 on NLVR2)
"""
import pytest
import torch

from oracle import trainer_oracle as to
from oracle import vilt_oracle as vo
from tests.golden_util import ALL_TASKS, TINY, TINY_HW, TINY_T, load
from tests.test_gpu_parity import _build
from tests.trainer_util import check_trajectory

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", list(to.SCENARIOS))
def test_reference_trainer_loops_on_the_cuda_path(tag):
    from climb_b200 import _lib
    from climb_b200.cl_algorithms import ExperienceReplayMemory
    from climb_b200.optim import ArenaAdamW
    dev = torch.device("cuda")
    sc = to.SCENARIOS[tag]
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=sc["seed"])
    learner = _build(TINY, ALL_TASKS, sd)
    pools, train_dl, val_dl, replay_dl = to.build_data(sc, TINY, TINY_T, TINY_HW)
    proc = to.PoolProcessor(pools, dev)
    learner.vilt_encoder.process_inputs = proc
    assert isinstance(learner.create_optimizer(sc["hparams"]), ArenaAdamW)
    launches0 = _lib.climb_launch_count()
    rec = to.run_scenario(sc, learner, dev, train_dl, val_dl, replay_dl, replay_memory_cls=ExperienceReplayMemory)
    assert _lib.climb_launch_count() - launches0 > 100, "the trajectory did not run on the CUDA kernels"
    g = load(tag)
    assert proc.calls == int(g["process_inputs_calls"])
    named_final = dict(learner.named_parameters())
    worst = check_trajectory(tag, rec, tol_loss=2e-2, tol_logits=5e-2, tol_update=1.5, tol_update_median=0.3,
                             named_final=named_final, named_init=sd, named_best=None,
                             replay_lr=sc["replay"]["hparams"]["lr"] if sc["replay"] else 0.0)
    print(tag, "losses", [round(x, 4) for x in rec["loss"]], "eval", rec["eval_score"], "worst update error", worst)
    # the deepcopy'd best model is a frozen snapshot on the device: evaluating it again reproduces its epoch
    assert to.reevaluate_snapshot(rec) <= 1e-6
    best = dict(rec["best_model"].named_parameters())
    if rec["best_epoch"] < sc["num_epochs"] - 1:
        moved = max((best[n] - named_final[n]).abs().max().item() for n in best)
        assert moved > 1e-4, "the snapshot followed the live parameters"
    # and its state dict has the reference's checkpoint keys (train_upstream_continual_learning.py:264-266)
    keys = set(rec["best_model"].state_dict().keys())
    assert {k for k in sd} <= keys
