"""The UNCHANGED harness under torchrun on real GPUs (run with 2 ranks):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_harness_check.py

Every rank runs the reference's own, unmodified trainer code (VQATrainer / NLVR2Trainer .train(), ExperienceReplayMemory --
imported from the staged archive, oracle/make_golden_trainer.run_reference_scenario) over the SAME seeded loaders, exactly as N
copies of train_upstream_continual_learning.py would. The only things that differ from the single-process run are what the
registry provides: the learner carries what create_continual_learner_map attaches under an initialised process group
(climb_b200.distributed.attach_if_distributed) -- its training forward runs on this rank's rows and returns the all-gathered
outputs of the whole batch, its backward ends with the gradient mean over ranks. Checked against the single-process golden
trajectories of the reference model (tests/golden/trainer_*.npz): per-step losses (the whole-batch loss, identical on every
rank), learning rates, evaluation logits and scores (evaluation stays replicated), final parameters -- and that all ranks end
with bit-identical parameters."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import ref_shim
    ref_shim.install()
    from oracle import trainer_oracle as to
    from oracle.make_golden_trainer import run_reference_scenario, scenario_state_dict
    from tests.golden_util import ALL_TASKS, TINY
    from tests.test_gpu_parity import _build
    from tests.trainer_util import check_trajectory
    from climb_b200 import distributed as cdist
    from climb_b200.modeling import model_configs

    ok = True
    from oracle.make_golden_trainer import prepare_adapters
    for tag in ("trainer_vqa_er", "trainer_nlvr2", "trainer_snli_ve", "trainer_snli_ve_ewc", "trainer_nlvr2_adapters"):
        sc = dict(to.SCENARIOS, **to.REFERENCE_SCENARIOS)[tag]
        assert sc["batch_size"] % world == 0
        sd, sd_full = scenario_state_dict(sc)
        learner = _build(TINY, ALL_TASKS, sd)
        ewc_cls = None
        if sc.get("adapters"):                                    # AdapterHandler, as the driver calls it before the task
            from climb_b200.cl_algorithms import AdapterHandler
            prepare_adapters(sc, learner, AdapterHandler)
            sd = sd_full
        if sc.get("ewc"):
            from climb_b200.cl_algorithms import EWC as ewc_cls
        sync = cdist.attach_if_distributed(learner)               # what create_*_continual_learner_model does
        assert sync is not None
        rec, extra = run_reference_scenario(tag, learner, str(dev), ewc_cls=ewc_cls,
                                            converter=model_configs["vilt-b200"]["batch2inputs_converter"])
        # (every rank computed the WHOLE batch's loss from the gathered logits: the recorded losses are already the golden's)
        t = torch.tensor(rec["loss"], dtype=torch.float64, device=dev)
        t0 = t.clone()
        dist.broadcast(t0, src=0)
        same_loss = bool(torch.equal(t, t0))
        worst = check_trajectory(tag, rec, tol_loss=2e-2, tol_logits=5e-2, tol_update=1.5, tol_update_median=0.3,
                                 named_final=dict(learner.named_parameters()), named_init=sd, named_best=None,
                                 replay_lr=sc["replay"]["hparams"]["lr"] if sc["replay"] else 0.0)
        flat = torch.cat([p.detach().flatten() for p in learner.parameters()])
        ref0 = flat.clone()
        dist.broadcast(ref0, src=0)
        same = bool(torch.equal(flat, ref0))
        if rank == 0:
            print(f"DDP-HARNESS {tag}: world {world}, losses (identical on every rank: {same_loss}) "
                  f"{[round(x, 4) for x in rec['loss']]} eval {rec['eval_score']} worst update error {worst}; parameters bit-identical "
                  f"across ranks: {same}", flush=True)
        ok = ok and same and same_loss
        sync.detach()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
