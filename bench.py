#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native ViLT hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5                      # our arm, 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W        # N ranks, one per GPU
    python bench.py --impl reference --steps 3 --warmup 1                # reference arithmetic on host cores

A "step" is one upstream-CL training step of CLiMB's sequential-FT VQA task on one synthetic batch
per GPU (BASELINE.json configs[1]): forward (B sequences of 40 text tokens + 197 image patches,
ViLT-base) -> BCEWithLogits x 3129 -> backward -> AdamW step (train_vqa.py:135-174), through the
public API of this repo (B200ViltContinualLearner + ArenaAdamW). One JSON line goes to stdout.

  value      samples/s over all GPUs, inputs already resident in HBM
  e2e        the same step fed from PINNED HOST buffers: the H2D copy of every step's inputs (prefetched
             on a copy stream, as a data loader would) and a D2H read of every step's loss are inside
             the timed region
  roofline   tcgen05 GEMM kernel: algorithmic FLOPs (2MNK per launch) / device time of those launches,
             measured with CUDA events on the launch stream inside this process (climb_profile_*),
             against the measured sustained bf16 peak of MEASURED_PEAKS.json; plus the attention
             kernels against the measured HBM copy bandwidth
  cpu_baseline  the UNMODIFIED reference (ViltContinualLearner + VQATrainer.train_step, imported from the archive that
             oracle/stage_ref.py staged at build() time; kind "reference") timed on this box's host cores on
             BASELINE config 1 (B=4) -- a reported baseline, not the target. Without a staged reference: the oracle
             port (kind "port")
  gpu_eager_baseline  the same unmodified reference module run eagerly on this GPU (fp32 and bf16 autocast) at the
             same batch size: the like-for-like GPU number
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_TEXT, IMG, N_LABELS = 40, 448, 3129
FLOPS_PER_SAMPLE_STEP = 128.88e9        # BASELINE.md section 3 (algorithmic, full-FT step)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")   # B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:          # region shorter than one sample: take what we have
            sm = [float(l.split(",")[1]) for _, l in self.lines[-2:] if l.count(",") >= 8] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference on host cores (oracle/ref_runner.py over the archive that
# oracle/stage_ref.py staged at build() time); the oracle port only if no reference was staged
# ----------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps: int, warmup: int, batch: int = 4):
    """BASELINE config 1: ViLT-base, seeded random init, VQA head, B=4 synthetic batch, fp32, all host threads; one
    fwd + loss + bwd + AdamW step per iteration. Returns (samples/s, cores, seconds/step, kind, what ran)."""
    from oracle import ref_runner
    if ref_runner.available():
        sps, cores, sec = ref_runner.cpu_reference(steps, warmup, batch)
        return sps, cores, sec, "reference", ("the unmodified reference: ViltContinualLearner + VQATrainer.train_step + create_optimizer "
                                              "(src/modeling/vilt.py, train_vqa.py:135-174) over the vendored transformers-4.17 ViltModel")
    from oracle import vilt_oracle as vo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dims = vo.ViltDims()
    sd = vo.synth_state_dict(dims, ["vqa"], seed=42)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=1e-2)
    b = vo.synth_batch("vqa", batch, dims, T=T_TEXT, image_hw=(IMG, IMG), seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, logits = vo.learner_forward(params, dims, "vqa", b)
        loss = vo.task_loss("vqa", logits, b["target"])
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return batch / sec, cores, sec, "port", "oracle/vilt_oracle.py (CPU restatement pinned to the reference by tests/golden); no staged reference found"


def workload_config(batch_per_gpu: int, world: int) -> dict:
    """The `config` object of the JSON line: the same for both arms (the reference arm adds its bounded sample)."""
    return {"workload": "ViLT-base sequential-FT VQAv2-shaped synthetic step (fwd+BCEx3129+bwd+AdamW), "
                        "BASELINE.json configs[1]", "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * world,
            "text_tokens": T_TEXT, "image": f"{IMG}x{IMG} -> 14x14 patches + cls", "seq_len": 237,
            "layers": 12, "hidden": 768, "parallelism": f"dp{world}",
            "l2": "activation working set ~5 GB per step >> 126 MB L2; 4 rotating input batches"}


REF_BATCH = 4       # BASELINE config 1: the reference's own CPU-runnable case = the bounded sample of the workload


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps / --warmup are honoured; each step is a bounded sample (B = 4 sequences, ~1 s on 16 cores), so that the driver's
    # K / W finish within a few minutes (the cap only guards against an accidental huge K)
    steps, warmup = max(1, min(args.steps, 200)), max(0, min(args.warmup, 50))
    sps, cores, sec, kind, what = cpu_reference_steps(steps, warmup, REF_BATCH)
    cfg = workload_config(REF_BATCH, 1)
    cfg["parallelism"] = "host cores (rank 0 only)"
    cfg["l2"] = "n/a (CPU)"
    cfg["sample"] = (f"each step = one B={REF_BATCH} batch of the same workload (BASELINE config 1: same model, sequence geometry, loss and "
                     "optimizer as the B=64-per-GPU arm), fp32 on the host cores")
    line = {
        "impl": "reference", "metric": "ViLT upstream-CL training throughput", "value": round(sps, 3), "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": round(sps, 3), "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} timed steps (median) of B={REF_BATCH}: {what}"},
        "e2e": {"value": round(sps, 3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def make_host_batch(B: int, seed: int, pin: bool):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 30000, (B, T_TEXT), generator=g)
    ids[:, 0], ids[:, -1] = 101, 102
    batch = {
        "input_ids": ids,
        "attention_mask": torch.ones(B, T_TEXT, dtype=torch.int64),
        "token_type_ids": torch.zeros(B, T_TEXT, dtype=torch.int64),
        "pixel_values": torch.rand(B, 3, IMG, IMG, generator=g) * 2 - 1,
    }
    tgt = torch.zeros(B, N_LABELS)
    for b in range(B):
        k = int(torch.randint(1, 4, (1,), generator=g))
        idx = torch.randperm(N_LABELS, generator=g)[:k]
        tgt[b, idx] = torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (k,), generator=g)]
    batch["target"] = tgt
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def run_ours(args):
    import torch.distributed as dist
    from climb_b200 import _lib, ops
    from climb_b200 import distributed as cdist
    from climb_b200.modeling import B200ViltConfig, B200ViltContinualLearner, B200ViltEncoderWrapper, B200ViltModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: climb_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        if args.nccl_ctas > 0:
            # NCCL's reduction kernels share the SMs with the persistent GEMM / attention grids: cap their CTAs (read by NCCL when
            # the communicator is created) and size the persistent grids to the SMs that are left while a reduction is in flight
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
        kw = {}
        if args.nccl_high_priority:
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
            kw["pg_options"] = opts
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180), **kw)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    B = args.batch
    tasks = ["vqa", "nlvr2", "snli-ve", "vcr"]
    task_specs = {"vqa": dict(num_labels=N_LABELS, num_images=1, model_type="classification"),
                  "nlvr2": dict(num_labels=2, num_images=2, model_type="classification"),
                  "snli-ve": dict(num_labels=3, num_images=1, model_type="classification"),
                  "vcr": dict(num_labels=4, num_images=1, model_type="multi-choice", num_choices=4)}
    torch.manual_seed(42)                       # same weights on every rank (train_upstream..py:103)
    learner = B200ViltContinualLearner(tasks, B200ViltEncoderWrapper(None, B200ViltModel(B200ViltConfig()), dev), 768,
                                       task_specs).to(dev)
    learner.train()
    if world > 1:
        # gradient all-reduce over NCCL from inside the chunked backward; the optimizer waits span by span (the plain training
        # step reads no gradient between backward() and step(): train_vqa.py:168-172)
        cdist.attach(learner, sm_reserve=int(os.environ.get("NCCL_MAX_CTAS", "0") or 0) if args.sm_reserve < 0 else args.sm_reserve,
                     defer_to_optimizer=bool(args.defer_optimizer), layers_per_chunk=args.layers_per_chunk, bucket_mb=args.bucket_mb)
    opt = learner.create_optimizer({"lr": 1e-4, "weight_decay": 1e-2, "adam_epsilon": 1e-8})
    total_steps = args.warmup * 2 + args.steps * 2 + 8
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: min(1.0, (s + 1) / 10.0) * max(0.0, 1.0 - s / (10.0 * total_steps)))

    n_batches = 4
    host = [make_host_batch(B, 1000 * rank + i, pin=True) for i in range(n_batches)]
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def step(batch):
        enc = {k: batch[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values")}
        _, logits = learner.forward_tensors("vqa", enc)
        loss = ops.vqa_loss(logits, batch["target"])
        loss.backward()
        opt.step()
        sched.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------------
    for i in range(args.warmup):
        step(resident[i % n_batches])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.climb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for i in range(args.steps):
        step(resident[i % n_batches])
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.climb_launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end: pinned host inputs, H2D prefetch on a copy stream, per-step D2H loss read ----
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    # a ring of RING slots: step i's inputs are uploaded while the previous RING - 1 steps may still be running, and the host
    # reads the loss of step i - (RING - 1). Every step's loss is read inside the timed region. RING = 2 (one step of
    # look-ahead) measured best: 3 changed nothing on 2 GPUs and cost 0.7 % on one.
    RING = max(2, args.e2e_ring)
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(RING)]
    loss_ev = [torch.cuda.Event() for _ in range(RING)]
    slots = [None] * RING
    slot_ready = [torch.cuda.Event() for _ in range(RING)]
    slot_free = [torch.cuda.Event() for _ in range(RING)]

    def upload(i):
        s = i % RING
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot_free[s])           # the step that last used this slot has finished
            slots[s] = {k: v.to(dev, non_blocking=True) for k, v in host[i % n_batches].items()}
            slot_ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in range(RING):
            slot_free[s].record(main_stream)
        for j in range(min(n, RING - 1)):
            upload(j)
        losses = []
        for i in range(n):
            s = i % RING
            if i + RING - 1 < n:
                upload(i + RING - 1)
            main_stream.wait_event(slot_ready[s])
            loss = step(slots[s])
            slot_free[s].record(main_stream)
            if i >= RING - 1:                              # read the loss of step i - (RING - 1): no pipeline bubble
                loss_ev[(i - RING + 1) % RING].synchronize()
                losses.append(float(loss_host[(i - RING + 1) % RING]))
            loss_host[s].copy_(loss.detach(), non_blocking=True)
            loss_ev[s].record(main_stream)
        for j in range(max(0, n - RING + 1), n):
            loss_ev[j % RING].synchronize()
            losses.append(float(loss_host[j % RING]))
        return losses

    e2e_loop(max(2, min(args.warmup, 3)))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    # ---- per-kernel device time for the roofline (events on the launch stream) -------------------
    # (every rank runs the two extra steps -- they contain the gradient all-reduce -- rank 0 reports)
    roofline, attn_roof = None, None
    barrier()
    _lib.check(_lib.climb_profile_begin())
    for i in range(2):
        step(resident[i % n_batches])
    prof = _lib.profile_end()
    barrier()
    if rank == 0:
        peaks = load_peaks()
        g_ms, g_flops, g_n = prof["gemm"]
        achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "gemm_pair_kernel<KIND> / gemm_pair_wgrad_kernel (cta_group::2) + gemm_fast_kernel<KIND> + gemm_bf16_tcgen05_kernel (all tcgen05 GEMM launches of the step)",
                    "achieved": round(achieved, 1),
                    "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": round(achieved / peaks["tf_sustained"], 4),
                    "traffic": traffic, "traffic_note": "mean DRAM bytes (ncu, read + write) over the nine hot GEMM launches of one layer, forward and "
                    "backward; per launch and per shape in profiles/gemm_traffic.json (measured <= algorithmic for every one of them)",
                    "peak_source": peaks["source"] + " sustained bf16 (kernel timed inside a long step)",
                    "launches_per_step": g_n // 2, "gemm_ms_per_step": round(g_ms / 2, 3),
                    "gemm_share_of_step": round((g_ms / 2) / (ms_total / args.steps), 3),
                    "note": "per-launch CUDA events serialise the stream, so these two profiling steps run without the "
                            "programmatic-dependent-launch overlap the timed region has"}
        af_ms, af_b, af_n = prof["attn_fwd"]
        ab_ms, ab_b, ab_n = prof["attn_bwd"]
        attn_roof = {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm"],
                     "fwd": {"achieved": round(af_b / (af_ms * 1e-3) / 1e9, 1) if af_ms else 0.0, "ms_per_step": round(af_ms / 2, 3)},
                     "bwd": {"achieved": round(ab_b / (ab_ms * 1e-3) / 1e9, 1) if ab_ms else 0.0, "ms_per_step": round(ab_ms / 2, 3)}}
        attn_roof["fwd"]["frac"] = round(attn_roof["fwd"]["achieved"] / peaks["hbm"], 4)
        attn_roof["bwd"]["frac"] = round(attn_roof["bwd"]["achieved"] / peaks["hbm"], 4)
        ap_ = os.path.join(ROOT, "profiles", "attention_traffic.json")
        if os.path.exists(ap_):            # ncu dram bytes per launch (one launch = B x heads of one layer)
            at = json.load(open(ap_))
            attn_roof["fwd"]["traffic"] = at["fwd"]["dram_bytes_per_launch"]
            attn_roof["bwd"]["traffic"] = at["bwd"]["dram_bytes_per_launch"]
            attn_roof["kernel"] = ("attn_tc_fwd3_kernel (P in tensor memory) / attn_tc_bwd2_kernel (persistent, tcgen05 + TMEM); achieved = algorithmic bytes / time; "
                                   "the instruction stream of the softmax / elementwise warps bounds these kernels before HBM does: DESIGN.md section 3")
        # the driver keeps `roofline` only: the attention kernels (the kernel BASELINE.json's metric names) ride inside it
        roofline["kernels"] = [
            {"name": "tcgen05 GEMMs (all launches)", "bound": "tensor", "achieved": roofline["achieved"], "peak": roofline["peak"],
             "unit": "TFLOP/s", "frac": roofline["frac"], "ms_per_step": roofline["gemm_ms_per_step"]},
            {"name": "attention forward (fused softmax(QK^T)V, per layer launch)", "bound": "hbm", "achieved": attn_roof["fwd"]["achieved"],
             "peak": peaks["hbm"], "unit": "GB/s", "frac": attn_roof["fwd"]["frac"], "ms_per_step": attn_roof["fwd"]["ms_per_step"],
             "algorithmic_bytes_per_launch": int(af_b / max(af_n, 1)), "traffic": attn_roof["fwd"].get("traffic")},
            {"name": "attention backward (dQ, dK, dV, per layer launch)", "bound": "hbm", "achieved": attn_roof["bwd"]["achieved"],
             "peak": peaks["hbm"], "unit": "GB/s", "frac": attn_roof["bwd"]["frac"], "ms_per_step": attn_roof["bwd"]["ms_per_step"],
             "algorithmic_bytes_per_launch": int(ab_b / max(ab_n, 1)), "traffic": attn_roof["bwd"].get("traffic")}]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, cores, sec, kind, what = cpu_reference_steps(steps=8, warmup=1)
        cpu_baseline = {"value": round(sps, 3), "unit": "samples/s", "cores": cores, "kind": kind,
                        "sample": f"8 timed steps (median) of B=4 fwd+loss+bwd+AdamW, BASELINE config 1, fp32: {what}"}

    # like-for-like GPU baseline: the unmodified reference module as-is on this GPU (eager PyTorch), same step, same batch size
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            from oracle import ref_runner
            if ref_runner.available():
                resident.clear()
                slots[:] = [None] * len(slots)
                torch.cuda.empty_cache()
                gpu_eager = ref_runner.gpu_eager(B)
                gpu_eager["what"] = ("the UNMODIFIED reference (ViltContinualLearner + VQATrainer.train_step + torch AdamW over the vendored "
                                     f"ViltModel) run eagerly on this GPU at B={B}; median of 3 steps after 2 warm-up steps")
        except Exception as e:          # a baseline must never take the product's line down
            gpu_eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    if rank == 0:
        peaks = load_peaks()
        line = {
            "metric": "ViLT upstream-CL training throughput", "value": round(value, 1), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_config(B, world), **({"ddp": f"gradient all-reduce (NCCL AVG, {args.bucket_mb:g} MB buckets) from inside the "
                           f"backward, issued in chunks of {args.layers_per_chunk} layers; AdamW "
                           + ("deferred span by span behind the in-flight reductions" if args.defer_optimizer else "after all of them")} if world > 1 else {})),
            "samples_per_s_per_gpu": round(value / world, 1),
            "model_tflops_per_gpu": round(value / world * FLOPS_PER_SAMPLE_STEP / 1e12, 1),
            "step_frac_of_tensor_roofline": round(value / world * FLOPS_PER_SAMPLE_STEP / 1e12 / peaks["tf_sustained"], 4),
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": round(e2e_ms / args.steps, 3),
                    "last_loss": round(losses[-1], 4)},
            "gpu_launches": int(launches),
            "roofline": roofline, "attention_roofline": attn_roof, "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager,
            "precision": precision_note(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def precision_note():
    """bf16 is the throughput mode; its measured error against the fp32 reference (tests/parity_gates.json holds 2 x these) and
    the precise mode's (bf16x3 split operands, fp32 activations: climb_b200 `precision='bf16x3'`) are reported beside the number."""
    p = os.path.join(ROOT, "profiles", "r2_parity_measured.json")
    if not os.path.exists(p):
        return None
    m = json.load(open(p))
    pick = lambda k: m.get(k)
    return {"mode": "bf16 operands, fp32 accumulate / residual stream / statistics",
            "measured_rel_error_vs_fp32_reference": {"bench_shape_b64_logits": pick("bench_shape_b64/logits"), "bench_shape_b64_pooled": pick("bench_shape_b64/pooled"),
                                                     "bench_shape_b64_worst_grad": pick("bench_shape_b64/grad"), "base_vqa_logits": pick("base_vqa/logits")},
            "precise_mode_rel_error": {"base_vqa_logits": pick("precise/base_vqa/logits"), "base_vqa_worst_grad": pick("precise/base_vqa/grad")},
            "source": "profiles/r2_parity_measured.json (pytest -m gpu -s on a B200, tools/update_gates.py)"}


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    # Everything else that might reach stdout (NCCL prints its version banner there when NCCL_DEBUG is set, library
    # chatter, warnings) is sent to stderr at the file-descriptor level, so stdout carries exactly one JSON line.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="sequences per GPU (the shipped scripts use 64)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nccl-ctas", type=int, default=0, help="N > 1 GPUs: cap on NCCL's CTAs (NCCL_MAX_CTAS, unless already set); 0 = NCCL's default")
    ap.add_argument("--sm-reserve", type=int, default=-1, help="SMs the persistent kernels leave free during a reduction; -1 = the NCCL CTA cap")
    ap.add_argument("--defer-optimizer", type=int, default=1, help="1: AdamW waits for the gradient all-reduce span by span (overlap); 0: after all of it")
    ap.add_argument("--layers-per-chunk", type=int, default=3)
    ap.add_argument("--nccl-high-priority", type=int, default=0, help="1: NCCL's kernels on a high-priority stream")
    ap.add_argument("--bucket-mb", type=float, default=64.0)
    ap.add_argument("--e2e-ring", type=int, default=2, help="input / loss slots of the end-to-end loop (look-ahead = ring - 1 steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-reference-on-this-GPU leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
