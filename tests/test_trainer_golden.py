"""CPU: the restated trainer loops (oracle/trainer_oracle.py) + the CPU oracle learner + THIS repo's
ExperienceReplayMemory replay the scenarios of tests/golden/trainer_*.npz, which were recorded from the
UNMODIFIED reference trainers (VQATrainer / NLVR2Trainer .train / .train_step / .eval, ExperienceReplayMemory,
ViltContinualLearner.create_optimizer, the polynomial-decay schedule). This pins
  * the restated loops that tests/test_gpu_zz_trainer.py then uses to drive the CUDA path, and
  * climb_b200.cl_algorithms.experience_replay: same memory indices, same replay batches, same task choices as
    the reference for the same python RNG seed (SURVEY.md section 8 row a21 -- host logic, runs without a GPU)."""
import numpy as np
import pytest
import torch

from oracle import trainer_oracle as to
from oracle import vilt_oracle as vo
from tests.golden_util import ALL_TASKS, TINY, TINY_HW, TINY_T, load
from tests.trainer_util import check_trajectory


@pytest.mark.parametrize("tag", list(to.SCENARIOS))
def test_restated_trainers_match_reference(tag):
    from climb_b200.cl_algorithms.experience_replay import ExperienceReplayMemory
    torch.manual_seed(0)
    sc = to.SCENARIOS[tag]
    sd = vo.synth_state_dict(TINY, ALL_TASKS, seed=sc["seed"])
    pools, train_dl, val_dl, replay_dl = to.build_data(sc, TINY, TINY_T, TINY_HW)
    proc = to.PoolProcessor(pools, torch.device("cpu"))
    learner = to.OracleLearner(TINY, ALL_TASKS, sd, proc)
    sampled = []
    mem_cls = ExperienceReplayMemory
    if sc["replay"]:
        class Recording(ExperienceReplayMemory):
            def run_replay_step(self, task_key, model):
                buf = self.memory_buffers[task_key]
                orig = buf.sample_replay_batch

                def sample():
                    b = orig()
                    sampled.append([h[1] for h in b["raw_texts"]])
                    return b
                buf.sample_replay_batch = sample
                try:
                    return super().run_replay_step(task_key, model)
                finally:
                    buf.sample_replay_batch = orig
        mem_cls = Recording
    rec = to.run_scenario(sc, learner, torch.device("cpu"), train_dl, val_dl, replay_dl, replay_memory_cls=mem_cls)
    g = load(tag)
    if sc["replay"]:
        assert np.array_equal(np.array(sampled), g["replay_samples"]), "replay batches differ from the reference's"
    assert proc.calls == int(g["process_inputs_calls"])
    named_final = dict(learner.named_parameters())
    named_best = dict(rec["best_model"].named_parameters())
    worst = check_trajectory(tag, rec, tol_loss=2e-4, tol_logits=2e-4, tol_update=2e-2, named_final=named_final,
                             named_init=sd, named_best=named_best,
                             replay_lr=sc["replay"]["hparams"]["lr"] if sc["replay"] else 0.0)
    print(tag, "worst final-parameter error relative to the update:", worst)
    assert to.reevaluate_snapshot(rec) == 0.0
    if rec["best_epoch"] < sc["num_epochs"] - 1:         # training went on after the snapshot: it must not have followed
        moved = max((named_best[n] - named_final[n]).abs().max().item() for n in named_best)
        assert moved > 1e-4


def test_memory_buffer_matches_reference_sampling():
    """TaskMemoryBuffer draws the memory with random.sample at construction and every replay batch with
    random.sample again (experience_replay.py:104-106, 118-122): the fixture holds the reference's draws."""
    import random
    import types
    from climb_b200.cl_algorithms.experience_replay import ExperienceReplayMemory
    tag = "trainer_vqa_er"
    sc, g = to.SCENARIOS[tag], load(tag)
    _, _, _, replay_dl = to.build_data(sc, TINY, TINY_T, TINY_HW)
    prev = to.TrainerOracle("nlvr2", replay_dl, replay_dl, sc["replay"]["hparams"], 1, torch.device("cpu"))
    random.seed(sc["seed"])
    mem = ExperienceReplayMemory()
    assert mem.do_replay() is False
    mem.add_task_memory_buffer(args=types.SimpleNamespace(batch_size=sc["batch_size"]), task_key="nlvr2",
                               task_config={"task_name": "nlvr2"}, task_trainer=prev,
                               memory_percentage=sc["replay"]["memory_percentage"], sampling_strategy="random")
    assert mem.do_replay() is True
    buf = mem.memory_buffers["nlvr2"]
    assert list(buf.memory_idxs) == list(g["memory_idxs"]) and len(buf) == len(g["memory_idxs"])
    assert buf.batch_size == sc["batch_size"] // 2            # NLVR2 pairs: half the batch (experience_replay.py:93-94)
    for want in g["replay_samples"]:
        assert mem.sample_replay_task() == "nlvr2"
        b = buf.sample_replay_batch()
        assert [h[1] for h in b["raw_texts"]] == list(want)
    with pytest.raises(AssertionError):
        mem.add_task_memory_buffer(args=types.SimpleNamespace(batch_size=4), task_key="vqa", task_config={"task_name": "vqa"},
                                   task_trainer=prev, memory_percentage=1.0, sampling_strategy="random")
    with pytest.raises((AssertionError, NotImplementedError)):
        mem.add_task_memory_buffer(args=types.SimpleNamespace(batch_size=4), task_key="vqa", task_config={"task_name": "vqa"},
                                   task_trainer=prev, memory_percentage=0.5, sampling_strategy="random-balanced")
