"""GPU: the reference's REAL trainer code drives the CUDA learner. tests/test_gpu_zz_trainer.py replays the trainer loops through
their restatement (oracle/trainer_oracle.py); here nothing is restated: the unmodified `VQATrainer.train()` /
`NLVR2Trainer.train()` / `SNLIVETrainer.train()` / `VCRTrainer.train()` (train_vqa.py:176-244, train_nlvr2.py, train_snli_ve.py,
train_vcr.py: all four task trainers) with their own `train_step`, `forward_pass`, `eval`,
`copy.deepcopy(model)` snapshots and polynomial-decay schedule, and the unmodified `ExperienceReplayMemory.run_replay_step`,
imported from the archive `build()` staged from /root/reference (oracle/_ref/, travels to the GPU box), call
`model(task_key=..., images=..., texts=...)`, `model.create_optimizer(...)` on B200ViltContinualLearner -- the drop-in claim of
INTEGRATION.md exercised by the harness's own code. Scenarios: sequential FT on NLVR2 / SNLI-VE / VCR, VQA with experience replay,
SNLI-VE with EWC (this repo's EWC class under the unmodified trainer; fixture from the reference's EWC) and NLVR2 with adapters (this
repo's AdapterHandler; fixture from the reference's handler + adapter-transformers). Each scenario runs in its own process (tests/ref_trainer_worker.py):
importing the reference puts its vendored transformers fork in front of the stock package.
Skipped when no reference is available (no /root/reference and no staged archive)."""
import os
import subprocess
import sys

import pytest

from oracle import trainer_oracle as to

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tag", list(to.SCENARIOS) + list(to.REFERENCE_SCENARIOS))
def test_unmodified_reference_trainers_drive_the_cuda_learner(tag):
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference not available: build() stages it from /root/reference into oracle/_ref/")
    r = subprocess.run([sys.executable, "-m", "tests.ref_trainer_worker", tag], cwd=ROOT, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    tail = "\n".join((r.stdout + "\n" + r.stderr).splitlines()[-25:])
    print(tail)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), tail
