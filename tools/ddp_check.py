"""Multi-GPU correctness check of the data-parallel path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py

Every rank takes its shard of one seeded global batch (climb_b200.distributed.shard_batch), runs forward + loss +
backward with the chunked, overlapped gradient all-reduce attached, and compares the averaged gradients it ends
up with against the gradients of the WHOLE batch computed locally without any communication (mean-reduced losses
over equal shards: the two must agree up to summation order). Also checks that all ranks hold bit-identical
gradients afterwards. Exit code 0 = pass."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import distributed as cdist, ops  # noqa: E402
from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    specs = {"vqa": dict(num_labels=3129, num_images=1, model_type="classification"),
             "nlvr2": dict(num_labels=2, num_images=2, model_type="classification")}
    cfg = B200ViltConfig(num_hidden_layers=4)          # ViLT-base width, 4 layers: two backward chunks + the embedding chunk
    torch.manual_seed(7)
    learner = B200ViltContinualLearner(["vqa", "nlvr2"], B200ViltEncoderWrapper(None, B200ViltModel(cfg), dev), 768, specs).to(dev)
    learner.train()
    g = torch.Generator().manual_seed(11)
    B = 4 * world
    ok = True
    for task, n_img in (("vqa", 1), ("nlvr2", 2)):
        ids = torch.randint(1000, 30000, (B, 40), generator=g)
        batch = {"input_ids": ids, "attention_mask": torch.ones(B, 40, dtype=torch.int64),
                 "token_type_ids": torch.zeros(B, 40, dtype=torch.int64),
                 "pixel_values": torch.rand(B * n_img, 3, 384, 384, generator=g) * 2 - 1}
        if task == "vqa":
            tgt = torch.zeros(B, 3129)
            tgt[torch.arange(B), torch.randint(0, 3129, (B,), generator=g)] = 1.0
        else:
            tgt = torch.randint(0, 2, (B,), generator=g)

        def run(b, t):
            enc = {k: v.to(dev) for k, v in b.items()}
            _, logits = learner.forward_tensors(task, enc)
            loss = ops.vqa_loss(logits, t.to(dev)) if task == "vqa" else ops.cross_entropy_loss(logits, t.to(dev))
            loss.backward()
            return {n: p.grad.detach().clone() for n, p in learner.named_parameters() if p.grad is not None}

        learner.zero_grad(set_to_none=True)
        full = run(batch, tgt)                                     # whole batch, no communication
        learner.zero_grad(set_to_none=True)
        sync = cdist.attach(learner, layers_per_chunk=2)
        text = {k: batch[k] for k in ("input_ids", "attention_mask", "token_type_ids")}
        shard = cdist.shard_batch(dict(text, target=tgt), rank, world)
        shard_px = cdist.shard_batch({"pixel_values": batch["pixel_values"]}, rank, world, group_size=n_img)
        t_shard = shard.pop("target")
        mine = run(dict(shard, **shard_px), t_shard)               # this rank's rows, gradients averaged over ranks
        sync.detach()
        worst = 0.0
        gscale = max(v.norm().item() for v in full.values())
        for n, ref in full.items():
            err = (mine[n] - ref).norm().item() / max(ref.norm().item(), 1e-3 * gscale)
            worst = max(worst, err)
        flat = torch.cat([v.flatten() for v in mine.values()])
        ref0 = flat.clone()
        dist.broadcast(ref0, src=0)
        same = bool(torch.equal(flat, ref0))
        if rank == 0:
            print(f"{task}: world {world}, worst relative gradient difference vs whole-batch {worst:.3e}, ranks identical: {same}", flush=True)
        ok = ok and worst < 2e-2 and same
    # ---- deferred exchange: ArenaAdamW waits span by span (GradSync(defer_to_optimizer=True)) + SM reserve: two training
    #      steps must leave every rank with bit-identical parameters (a span consumed before its all-reduce completed would
    #      hold this rank's own gradient) and agree with the plain finish()-waits-for-everything path up to the run-to-run
    #      noise of the fp32 atomics in the weight gradients ----
    import copy
    sd0 = copy.deepcopy(learner.state_dict())
    ids = torch.randint(1000, 30000, (4, 40), generator=torch.Generator().manual_seed(100 + rank))
    enc = {"input_ids": ids.to(dev), "attention_mask": torch.ones(4, 40, dtype=torch.int64, device=dev),
           "token_type_ids": torch.zeros(4, 40, dtype=torch.int64, device=dev),
           "pixel_values": (torch.rand(4, 3, 384, 384, generator=torch.Generator().manual_seed(200 + rank)) * 2 - 1).to(dev)}
    tgt = torch.zeros(4, 3129, device=dev)
    tgt[torch.arange(4), torch.arange(4) * 7 + rank] = 1.0
    finals = []
    for defer, reserve in ((False, 0), (True, 8)):
        learner.load_state_dict(sd0)
        learner.zero_grad(set_to_none=True)
        sync = cdist.attach(learner, layers_per_chunk=2, bucket_mb=8.0, defer_to_optimizer=defer, sm_reserve=reserve)
        opt = learner.create_optimizer({"lr": 1e-3, "weight_decay": 1e-2, "adam_epsilon": 1e-8})
        for _ in range(2):
            _, logits = learner.forward_tensors("vqa", enc)
            ops.vqa_loss(logits, tgt).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        sync.detach()
        finals.append(torch.cat([p.detach().flatten() for p in learner.parameters()]).clone())
    from climb_b200 import _lib
    p0 = torch.cat([sd0[n].flatten().float() for n, _ in learner.named_parameters()])
    upd = (finals[0] - p0).norm().item()
    diff = (finals[1] - finals[0]).norm().item() / max(upd, 1e-30)
    other = finals[1].clone()
    dist.broadcast(other, src=0)
    ranks_same = bool(torch.equal(other, finals[1]))
    reserve_reset = _lib.climb_set_sm_reserve(-1) == 0
    if rank == 0:
        print(f"deferred per-span optimizer + SM reserve: ranks bit-identical after 2 steps: {ranks_same}; difference to the plain path "
              f"/ size of the update: {diff:.3e}; reserve released: {reserve_reset}", flush=True)
    ok = ok and ranks_same and diff < 2e-2 and reserve_reset
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
