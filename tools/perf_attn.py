"""Attention forward / backward timing at the bench shape (CUDA events, warm-up; ROT rotating input sets so that
every launch reads its operands from HBM rather than the 126 MB L2). Dev tool: CLIMB_ATTN_V1=1 for the A/B."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from climb_b200 import _lib as L  # noqa: E402


def main():
    B, Lq, H = int(os.environ.get("B", 64)), int(os.environ.get("LQ", 237)), 12
    rot = int(os.environ.get("ROT", 6))
    d = H * 64
    dev = "cuda"
    qkvs = [(torch.randn(B, Lq, 3 * d, device=dev)).bfloat16() for _ in range(rot)]
    kb = torch.zeros(B, Lq, device=dev)
    outs = [L.attention_fwd(q, kb, B, Lq, H, 0.125) for q in qkvs]
    dctxs = [torch.randn(B, Lq, d, device=dev).bfloat16() for _ in range(rot)]
    torch.cuda.synchronize()

    def timeit(fn, iters=30, warm=6):
        for i in range(warm):
            fn(i % rot)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(iters):
            fn(i % rot)
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    ms_f = timeit(lambda i: L.attention_fwd(qkvs[i], kb, B, Lq, H, 0.125))
    ms_b = timeit(lambda i: L.attention_bwd(qkvs[i], kb, outs[i][0], dctxs[i], outs[i][1], B, Lq, H, 0.125))
    cs = torch.zeros(3 * d, device=dev)
    ms_bc = timeit(lambda i: L.attention_bwd(qkvs[i], kb, outs[i][0], dctxs[i], outs[i][1], B, Lq, H, 0.125, colsum=cs))
    by_f = B * (4 * Lq * d * 2 + Lq * H * 4)
    print(f"attn B={B} L={Lq} v1={os.environ.get('CLIMB_ATTN_V1', '0')}: fwd {ms_f * 1e3:.1f} us ({by_f / ms_f / 1e6:.0f} GB/s) "
          f"bwd {ms_b * 1e3:.1f} us ({2 * by_f / ms_b / 1e6:.0f} GB/s) bwd+bias-grads {ms_bc * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
