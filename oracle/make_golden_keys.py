"""Generate tests/golden/state_dict_keys.json: the checkpoint surface of the UNMODIFIED reference learners -- every
state_dict key with shape and dtype, and the named_parameters() order -- for the objects CLiMB's driver saves and reloads
(train_upstream_continual_learning.py:264-266: best_model.state_dict() and best_model.get_encoder().state_dict();
eval_forgetting: model.load_state_dict(torch.load(path)), train_vqa.py:269-282) and that EWC / the optimizers iterate.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_keys

TEST INFRASTRUCTURE (build container only). Scenarios, all on the tiny geometry of the other fixtures:
  vilt            ViltContinualLearner over the four tasks
  vilt_adapters   the same after AdapterHandler-style add_adapter('vqa' / 'nlvr2', Houlsby rf 4) + train_adapter('nlvr2')
                  (also records requires_grad per parameter)
  vilt_pfeiffer   one Pfeiffer adapter (output site only)
  viltbert        ViltBertContinualLearner (viltbert_encoder.vilt.* / viltbert_encoder.bert.*)
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.make_golden import ALL_TASKS, GOLDEN_DIR, TINY, TINY_BERT, build_reference_learner, build_reference_viltbert  # noqa: E402
from oracle.vilt_oracle import synth_state_dict, synth_viltbert_state_dict  # noqa: E402


def describe(model):
    sd = model.state_dict()
    enc = model.get_encoder().state_dict()
    return {
        "state_dict": [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()],
        "encoder_state_dict": [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in enc.items()],
        "named_parameters": [[n, bool(p.requires_grad)] for n, p in model.named_parameters()],
    }


def main():
    ref_shim.install()
    from transformers.adapters import AdapterConfig
    out = {}
    learner = build_reference_learner(TINY, ALL_TASKS, synth_state_dict(TINY, ALL_TASKS, seed=1))
    out["vilt"] = describe(learner)
    cfg = AdapterConfig.load("houlsby").to_dict()
    cfg["reduction_factor"] = 4
    for task in ("vqa", "nlvr2"):
        learner.add_adapter(task, config=AdapterConfig.from_dict(cfg))
    learner.train_adapter("nlvr2")
    learner.set_active_adapters("nlvr2")
    out["vilt_adapters"] = describe(learner)
    learner2 = build_reference_learner(TINY, ALL_TASKS, synth_state_dict(TINY, ALL_TASKS, seed=1))
    pf = AdapterConfig.load("pfeiffer").to_dict()
    pf["reduction_factor"] = 2
    learner2.add_adapter("snli-ve", config=AdapterConfig.from_dict(pf))
    out["vilt_pfeiffer"] = describe(learner2)
    vb = build_reference_viltbert(TINY, TINY_BERT, ALL_TASKS, synth_viltbert_state_dict(TINY, TINY_BERT, ALL_TASKS, seed=1))
    out["viltbert"] = describe(vb)
    with open(os.path.join(GOLDEN_DIR, "state_dict_keys.json"), "w") as f:
        json.dump(out, f, indent=0)
    for k, v in out.items():
        print(k, len(v["state_dict"]), "state_dict keys,", len(v["encoder_state_dict"]), "encoder keys,", len(v["named_parameters"]), "parameters,",
              sum(1 for _, rg in v["named_parameters"] if rg), "trainable")


if __name__ == "__main__":
    main()
