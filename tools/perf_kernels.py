"""Per-kernel timing on the GPU box (CUDA events, warm-up, inputs >> L2 are not needed here: each
GEMM's working set at B=64 already exceeds what stays hot between different calls). Dev tool."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    B = int(os.environ.get("B", 64))
    Lq, d, ff, H = 237, 768, 3072, 12
    M = B * Lq
    dev = "cuda"
    out = {"B": B, "M": M}
    bf = lambda *s: (torch.randn(*s, device=dev) * 0.05).bfloat16()
    x, w_qkv, w_o, w1, w2 = bf(M, d), bf(3 * d, d), bf(d, d), bf(ff, d), bf(d, ff)
    inter, dy_d, dy_ff, dy_qkv = bf(M, ff), bf(M, d), bf(M, ff), bf(M, 3 * d)
    bias_ff, bias_d, bias_qkv = torch.randn(ff, device=dev), torch.randn(d, device=dev), torch.randn(3 * d, device=dev)
    res = torch.randn(M, d, device=dev)
    o_qkv = torch.empty(M, 3 * d, device=dev, dtype=torch.bfloat16)
    o_ff = torch.empty(M, ff, device=dev, dtype=torch.bfloat16)
    aux_ff = torch.empty(M, ff, device=dev, dtype=torch.bfloat16)
    o_d32 = torch.empty(M, d, device=dev)
    o_d16 = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    gw_qkv, gw_o, gw1, gw2 = torch.zeros(3 * d, d, device=dev), torch.zeros(d, d, device=dev), torch.zeros(ff, d, device=dev), torch.zeros(d, ff, device=dev)
    cases = {
        "fwd_qkv": (lambda bn: L.gemm(x, w_qkv, o_qkv, bias=bias_qkv, block_n=bn), 2 * M * 3 * d * d),
        "fwd_o_res": (lambda bn: L.gemm(x, w_o, o_d32, bias=bias_d, residual=res, block_n=bn), 2 * M * d * d),
        "fwd_fc1_gelu": (lambda bn: L.gemm(x, w1, o_ff, bias=bias_ff, epilogue=L.EPI_GELU_SAVE_GRAD, aux=aux_ff, block_n=bn), 2 * M * ff * d),
        "fwd_fc2_res": (lambda bn: L.gemm(inter, w2, o_d32, bias=bias_d, residual=res, block_n=bn), 2 * M * ff * d),
        "dgrad_fc2_mulaux": (lambda bn: L.gemm(dy_d, w2, o_ff, b_mn_major=True, epilogue=L.EPI_MUL_AUX, aux=aux_ff, M=M, N=ff, K=d, block_n=bn), 2 * M * ff * d),
        "dgrad_fc1": (lambda bn: L.gemm(dy_ff, w1, o_d16, b_mn_major=True, M=M, N=d, K=ff, block_n=bn), 2 * M * ff * d),
        "dgrad_qkv": (lambda bn: L.gemm(dy_qkv, w_qkv, o_d16, b_mn_major=True, M=M, N=d, K=3 * d, block_n=bn), 2 * M * 3 * d * d),
        "dgrad_o": (lambda bn: L.gemm(dy_d, w_o, o_d16, b_mn_major=True, M=M, N=d, K=d, block_n=bn), 2 * M * d * d),
        "wgrad_fc2": (lambda bn: L.gemm(dy_d, inter, gw2, a_mn_major=True, b_mn_major=True, accumulate=True, M=d, N=ff, K=M, block_n=bn), 2 * M * ff * d),
        "wgrad_fc1": (lambda bn: L.gemm(dy_ff, x, gw1, a_mn_major=True, b_mn_major=True, accumulate=True, M=ff, N=d, K=M, block_n=bn), 2 * M * ff * d),
        "wgrad_qkv": (lambda bn: L.gemm(dy_qkv, x, gw_qkv, a_mn_major=True, b_mn_major=True, accumulate=True, M=3 * d, N=d, K=M, block_n=bn), 2 * M * 3 * d * d),
        "wgrad_o": (lambda bn: L.gemm(dy_d, x, gw_o, a_mn_major=True, b_mn_major=True, accumulate=True, M=d, N=d, K=M, block_n=bn), 2 * M * d * d),
    }
    for name, (fn, flops) in cases.items():
        for bn in (0,):
            ms = timeit(lambda: fn(bn))
            out[f"{name}/bn{bn}"] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}
            print(name, bn, out[f"{name}/bn{bn}"], flush=True)
    if os.environ.get("ONLY_GEMM") == "1":          # A/B runs of the GEMM kernels only (e.g. CLIMB_GEMM_PAIR=0/1)
        return
    # cuBLAS reference points (library, for context only)
    ms = timeit(lambda: torch.matmul(x, w1.t()))
    out["cublas_fc1"] = {"ms": round(ms, 4), "tflops": round(2 * M * ff * d / ms / 1e9, 1)}
    ms = timeit(lambda: torch.matmul(inter, w2.t()))
    out["cublas_fc2"] = {"ms": round(ms, 4), "tflops": round(2 * M * ff * d / ms / 1e9, 1)}
    print("cublas", out["cublas_fc1"], out["cublas_fc2"], flush=True)
    # attention
    qkv = bf(B, Lq, 3 * d) * 10
    kb = torch.zeros(B, Lq, device=dev)
    ms_f = timeit(lambda: L.attention_fwd(qkv, kb, B, Lq, H, 0.125))
    ctx, lse = L.attention_fwd(qkv, kb, B, Lq, H, 0.125)
    dctx = bf(B, Lq, d)
    ms_b = timeit(lambda: L.attention_bwd(qkv, kb, ctx, dctx, lse, B, Lq, H, 0.125))
    fl = 4 * B * H * Lq * Lq * 64
    by_f = B * (4 * Lq * d * 2 + Lq * H * 4)
    out["attn_fwd"] = {"ms": round(ms_f, 4), "tflops": round(fl / ms_f / 1e9, 1), "GBs": round(by_f / ms_f / 1e6, 1)}
    out["attn_bwd"] = {"ms": round(ms_b, 4), "tflops": round(2 * fl / ms_b / 1e9, 1), "GBs": round(2 * by_f / ms_b / 1e6, 1)}
    print("attn", out["attn_fwd"], out["attn_bwd"], flush=True)
    # layernorm
    xf = torch.randn(M, d, device=dev)
    g, b_ = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    ms = timeit(lambda: L.layernorm_fwd(xf, g, b_, 1e-12))
    out["ln_fwd"] = {"ms": round(ms, 4), "GBs": round(M * d * 6 / ms / 1e6, 1)}
    yb, _, mean, rstd = L.layernorm_fwd(xf, g, b_, 1e-12)
    dx, dxb = torch.empty_like(xf), torch.empty(M, d, device=dev, dtype=torch.bfloat16)
    dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ms = timeit(lambda: L.layernorm_bwd(yb, xf, g, b_, mean, rstd, dres=xf, dx_f32=dx, dx_bf16=dxb, dgamma=dg, dbeta=db))
    out["ln_bwd"] = {"ms": round(ms, 4), "GBs": round(M * d * (2 + 4 + 4 + 4 + 2) / ms / 1e6, 1)}
    cs = torch.zeros(ff, device=dev)
    ms = timeit(lambda: L.colsum(dy_ff, cs))
    out["colsum_ff"] = {"ms": round(ms, 4), "GBs": round(M * ff * 2 / ms / 1e6, 1)}
    print("ln/colsum", out["ln_fwd"], out["ln_bwd"], out["colsum_ff"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    tag = "_generic" if os.environ.get("CLIMB_GEMM_GENERIC") == "1" else ""
    json.dump(out, open(f"gpurun_out/perf_kernels{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
