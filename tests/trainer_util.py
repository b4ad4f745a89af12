"""Shared checks for tests/golden/trainer_*.npz (written by oracle/make_golden_trainer.py from the UNMODIFIED
reference trainers): compare a replay of the scenario -- on the CPU oracle or on the CUDA path -- with what the
reference recorded."""
import numpy as np
import torch

from tests.golden_util import grad_sample_index, load


def check_trajectory(tag, rec, tol_loss, tol_logits, tol_update, named_final, named_init, named_best, replay_lr=0.0,
                     tol_update_median=None):
    """rec: oracle.trainer_oracle.run_scenario's record.
      * learning rates: exact (host arithmetic);
      * per-step training losses and replay losses: relative error <= tol_loss;
      * evaluation after every epoch: logits within tol_logits (relative Frobenius) and the SCORE equal to the
        reference's unless a validation sample's decision is a near tie in the reference itself (top-2 margin
        below the logit error bound) -- each such sample may move the score by its own share only;
      * best epoch / best score: equal when no near tie is involved;
      * final parameters: ||theta - theta_ref|| <= tol_update * ||theta_ref - theta_init|| per tensor (error
        relative to the UPDATE the run made, not to the parameter), on the stored sample; optionally the median
        over tensors <= tol_update_median (Adam turns small gradients into sign-like steps, so tensors that only
        see a few noisy steps -- the replayed task's head -- carry large relative errors in any precision);
      * the deepcopy'd best model really is the snapshot of its epoch: parameter norms match the reference's."""
    g = load(tag)
    assert np.allclose(rec["lr"], g["lr"], rtol=1e-12, atol=0), (rec["lr"], g["lr"])
    loss, ref_loss = np.array(rec["loss"]), g["loss"]
    assert loss.shape == ref_loss.shape
    rel = np.abs(loss - ref_loss) / np.abs(ref_loss)
    assert rel.max() <= tol_loss, ("training loss", rel, loss, ref_loss)
    if "replay_loss" in g.files:
        rl = np.array([l for _, l in rec["replay"]])
        assert [t for t, _ in rec["replay"]] == list(g["replay_task"])
        # a replay loss can sit near zero-crossing of nothing: CE >= 0 and of order 1 here, relative is fine
        assert (np.abs(rl - g["replay_loss"]) / np.abs(g["replay_loss"])).max() <= 2 * tol_loss, ("replay loss", rl, g["replay_loss"])
    n_val = g["eval_logits/0"].shape[0]
    ties_any = False
    for e, score in enumerate(rec["eval_score"]):
        ref = torch.from_numpy(g[f"eval_logits/{e}"]).float()
        got = rec["eval_logits"][e].float().reshape(ref.shape)
        err = ((got - ref).norm() / ref.norm()).item()
        assert err <= tol_logits, (f"eval logits epoch {e}", err)
        top2 = ref.topk(2, dim=-1).values
        margin = (top2[:, 0] - top2[:, 1])
        bound = 2 * (got - ref).abs().max().item()
        ties = int((margin <= bound).sum())
        ties_any |= ties > 0
        agree = (got.argmax(-1) == ref.argmax(-1)) | (margin <= bound)
        assert bool(agree.all()), (f"arg-max decisions differ beyond the near ties, epoch {e}", got.argmax(-1), ref.argmax(-1))
        assert abs(score - float(g["eval_score"][e])) <= ties * 100.0 / n_val + 1e-6, (e, score, float(g["eval_score"][e]), ties)
    if not ties_any:
        assert rec["best_epoch"] == int(g["best_epoch"])
        assert abs(rec["best_score"] - float(g["best_score"])) < 1e-6
    checked = 0
    worst = (0.0, None)
    all_err = []
    lr_budget = float(np.sum(g["lr"])) + (float(len(g["replay_loss"])) * replay_lr if "replay_loss" in g.files else 0.0)
    for key in g.files:
        if not key.startswith("final_sample/"):
            continue
        name = key[len("final_sample/"):]
        ref = torch.from_numpy(g[key]).double()
        idx = torch.from_numpy(grad_sample_index(named_final[name].numel())) if ref.numel() < named_final[name].numel() else None
        pick = (lambda t: t.detach().double().cpu().flatten()[idx]) if idx is not None else (lambda t: t.detach().double().cpu().flatten())
        got, init = pick(named_final[name]), pick(named_init[name])
        upd = (ref - init).norm().item()
        if name.endswith("attention.key.bias") or name == "task_layer.vcr.1.bias":
            # (the multi-choice head's bias adds the same constant to the four choice logits of a sample: the cross entropy
            #  over the choices does not see it either)
            # analytically zero gradient (softmax is invariant to a shift of all keys): what reaches Adam is rounding
            # noise, which Adam normalises into +-lr steps -- in the reference as much as anywhere else. Only bounded.
            assert (got - init).abs().max().item() <= 1.05 * lr_budget, name
            continue
        if upd == 0.0:                      # parameter off the path (e.g. another task's head): must be untouched
            assert (got - init).abs().max().item() == 0.0, name
            continue
        err = (got - ref).norm().item() / upd
        worst = max(worst, (err, name))
        all_err.append(err)
        assert err <= tol_update, (name, err)
        checked += 1
    assert checked > 20, checked
    if tol_update_median is not None:
        assert float(np.median(all_err)) <= tol_update_median, float(np.median(all_err))
    if not ties_any and named_best is not None:
        for key in g.files:
            if key.startswith("best_norm/"):
                name = key[len("best_norm/"):]
                ref_n = float(g[key])
                fin_n = float(g["final_norm/" + name])
                got_n = named_best[name].detach().double().norm().item()
                # the snapshot must be the best epoch's parameters, not the final ones, wherever the two differ measurably
                assert abs(got_n - ref_n) <= max(0.25 * abs(fin_n - ref_n), 5e-3 * ref_n), (name, got_n, ref_n, fin_n)
    return worst
