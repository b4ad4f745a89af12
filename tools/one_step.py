"""One steady-state training step of the bench workload inside an NVTX range, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches.csv python tools/one_step.py
Dev / profiling tool; numbers printed under ncu are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from climb_b200 import ops  # noqa: E402
from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel  # noqa: E402

B = int(os.environ.get("B", 64))
MODE = os.environ.get("MODE", "vqa")          # vqa (bench workload) | adapters (BASELINE config 3: NLVR2 pairs, Houlsby rf 16)
WARM = int(os.environ.get("WARM", 3))
dev = torch.device("cuda")
specs = {"vqa": dict(num_labels=3129, num_images=1, model_type="classification"),
         "nlvr2": dict(num_labels=2, num_images=2, model_type="classification")}
torch.manual_seed(42)
tasks = ["vqa"] if MODE == "vqa" else ["vqa", "nlvr2"]
learner = B200ViltContinualLearner(tasks, B200ViltEncoderWrapper(None, B200ViltModel(B200ViltConfig()), dev), 768, specs).to(dev)
learner.train()
if MODE == "adapters":
    learner.add_adapter("nlvr2", "houlsby")
    learner.train_adapter("nlvr2")
    learner.set_active_adapters("nlvr2")
opt = learner.create_optimizer({"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8})
batch = {k: v.to(dev) for k, v in bench.make_host_batch(B, 0, pin=False).items()}
if MODE == "adapters":
    batch = {k: (v[:B // 2] if k != "pixel_values" else v) for k, v in batch.items()}       # B/2 texts, B images (pairs)
    batch["target"] = torch.randint(0, 2, (B // 2,), device=dev)


def step():
    enc = {k: batch[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values")}
    if MODE == "adapters":
        _, logits = learner.forward_tensors("nlvr2", enc)
        loss = ops.cross_entropy_loss(logits, batch["target"])
    else:
        _, logits = learner.forward_tensors("vqa", enc)
        loss = ops.vqa_loss(logits, batch["target"])
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


for _ in range(WARM):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: capture exactly this step (all threads)
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
