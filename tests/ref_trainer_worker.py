"""Worker of tests/test_gpu_zzz_reference_trainer.py (own process: the reference's vendored transformers 4.17 shadows the stock
package once oracle.ref_shim is installed):

    python -m tests.ref_trainer_worker <scenario tag>

The UNMODIFIED reference trainers (VQATrainer / NLVR2Trainer .train / .train_step / .eval, train_vqa.py / train_nlvr2.py) and the
UNMODIFIED ExperienceReplayMemory (experience_replay.py), imported from /root/reference or from the archive build() staged,
drive the CUDA learner; the trajectory is held to the fixtures the same code recorded on the reference model."""
import sys

import torch


def main(tag):
    from oracle import ref_shim
    ref_shim.install()
    from oracle import trainer_oracle as to
    from oracle import vilt_oracle as vo
    from oracle.make_golden_trainer import prepare_adapters, run_reference_scenario, scenario_state_dict
    from tests.golden_util import ALL_TASKS, TINY, load
    from tests.test_gpu_parity import _build
    from tests.trainer_util import check_trajectory
    from climb_b200 import _lib
    from climb_b200.optim import ArenaAdamW

    sc = dict(to.SCENARIOS, **to.REFERENCE_SCENARIOS)[tag]
    sd, sd_full = scenario_state_dict(sc)
    if sc.get("encoder") == "viltbert":
        from oracle.make_golden import TINY_BERT
        from tests.test_gpu_viltbert import _build as _build_viltbert
        learner = _build_viltbert(TINY, TINY_BERT, ALL_TASKS, sd)
    else:
        learner = _build(TINY, ALL_TASKS, sd)
    if sc.get("adapters"):
        from climb_b200.cl_algorithms import AdapterHandler       # this repo's handler behind the reference's call surface
        prepare_adapters(sc, learner, AdapterHandler)
        sd = sd_full
        g0 = load(tag)
        assert sorted(n for n, p in learner.named_parameters() if p.requires_grad) == list(g0["trainable"])
    assert isinstance(learner.create_optimizer(sc["hparams"]), ArenaAdamW)
    launches0 = _lib.climb_launch_count()
    ewc_cls = None
    if sc.get("ewc"):
        from climb_b200.cl_algorithms import EWC as ewc_cls        # device-resident theta*, F behind the reference's hook surface
    rec, extra = run_reference_scenario(tag, learner, "cuda", ewc_cls=ewc_cls)
    assert _lib.climb_launch_count() - launches0 > 100, "the trajectory did not run on the CUDA kernels"
    import train.visionlanguage_tasks.train_vqa as tv
    assert "reference" in tv.__file__ or "climb_b200_reference" in tv.__file__, tv.__file__      # the reference's own module ran
    g = load(tag)
    assert extra["proc"].calls == int(g["process_inputs_calls"])
    if sc["replay"]:
        # python's RNG drives the memory contents and the sampled replay batches: identical to the reference run
        assert extra["memory_idxs"] == list(g["memory_idxs"])
        assert extra["sampled"] == g["replay_samples"].tolist()
    if sc.get("ewc"):
        import numpy as np
        got = np.array([l for _, l in rec["ewc"]])
        assert [t for t, _ in rec["ewc"]] == list(g["ewc_task"])
        ref = g["ewc_loss"]
        assert got.shape == ref.shape and got[0] == 0.0 and ref[0] == 0.0          # first step: theta == theta*
        # the penalty is lambda * sum F (theta - theta*)^2 after a few Adam steps: quadratic in parameter differences that
        # bf16 gradients move by a few percent -- a looser bound than the task loss, same reasoning as the replay losses
        rel = np.abs(got[1:] - ref[1:]) / np.abs(ref[1:])
        print("EWC penalties", got.round(4).tolist(), "reference", ref.round(4).tolist(), "rel", rel.round(3).tolist())
        assert rel.max() <= 0.03, rel          # measured on a B200: 5e-3
    worst = check_trajectory(tag, rec, tol_loss=2e-2, tol_logits=5e-2, tol_update=1.5, tol_update_median=0.3,
                             named_final=dict(learner.named_parameters()), named_init=sd, named_best=None,
                             replay_lr=sc["replay"]["hparams"]["lr"] if sc["replay"] else 0.0)
    print(f"REFERENCE-TRAINER {tag}: losses {[round(float(x), 4) for x in rec['loss']]} eval {rec['eval_score']} best epoch "
          f"{rec['best_epoch']} worst update error {worst}")
    assert isinstance(rec["best_model"], type(learner))         # copy.deepcopy(model) inside the reference's train()
    torch.cuda.synchronize()
    print("OK")


if __name__ == "__main__":
    main(sys.argv[1])
