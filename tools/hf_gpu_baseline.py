"""Like-for-like GPU context number: the SAME training step as bench.py -- ViLT-base, B sequences
of 40 text tokens + 448 x 448 image, VQA head, BCEWithLogits x 3129, AdamW -- run as plain eager PyTorch on the
same B200 with the stock `transformers` ViltModel that ships in this image (the reference vendors the 4.17
modeling_vilt.py of the same lineage, which cannot travel to the GPU box; this is an informal stand-in for
"the reference module run as-is on the GPU", not the reference and not a parity oracle).

    python tools/hf_gpu_baseline.py [--batch 64] [--steps 5]

Prints one JSON object per mode: fp32 (TF32 matmuls off, the reference's arithmetic), tf32, bf16 autocast."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    from transformers import ViltConfig, ViltModel
    dev = torch.device("cuda")
    out = []
    for mode in ("fp32", "tf32", "bf16-autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        torch.manual_seed(42)
        enc = ViltModel(ViltConfig()).to(dev).train()
        head = nn.Sequential(nn.Linear(768, 1536), nn.LayerNorm(1536), nn.GELU(), nn.Linear(1536, 3129)).to(dev)
        params = list(enc.parameters()) + list(head.parameters())
        opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
        batch = {k: v.to(dev) for k, v in bench.make_host_batch(a.batch, 0, pin=False).items()}
        pm = torch.ones(a.batch, bench.IMG, bench.IMG, dtype=torch.long, device=dev)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16-autocast")):
                pooled = enc(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                             token_type_ids=batch["token_type_ids"], pixel_values=batch["pixel_values"], pixel_mask=pm).pooler_output
                logits = head(pooled)
            loss = nn.functional.binary_cross_entropy_with_logits(logits.float(), batch["target"]) * 3129
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss

        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        line = {"impl": "stock transformers ViltModel, eager PyTorch, same B200", "mode": mode, "batch": a.batch,
                "ms_per_step": round(ms, 2), "samples_per_s": round(a.batch / ms * 1e3, 1), "loss": round(float(loss), 3),
                "torch": torch.__version__}
        out.append(line)
        print(json.dumps(line), flush=True)
        del enc, head, opt, params
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/hf_gpu_baseline.json", "w"), indent=1)


if __name__ == "__main__":
    main()
