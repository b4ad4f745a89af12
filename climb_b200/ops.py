"""Autograd functions over the C ABI for the pieces of the hot path that live OUTSIDE ViltModel:
the task heads (src/modeling/vilt.py:179-203), the trainers' losses (train_vqa.py:95,157;
train_nlvr2.py:80,133) and EWC's penalty (src/cl_algorithms/ewc.py:75-87).

Every forward / backward here launches kernels from libclimb_b200.so; none falls back to torch math.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib


# ---- arithmetic mode (include/climb_b200.h: CLIMB_PREC_*) ---------------------------------------------------------------
# "bf16"   : bf16 tensor-core operands, fp32 accumulation / residual stream / statistics: the throughput mode (default).
# "bf16x3" : the parity gate -- every contraction as three accumulating tcgen05 launches over split operands
#            (x = hi + lo, both bf16), fp32 activations in between, fp32 attention: logits within 1e-3 of the fp32
#            reference (the north star's tolerance), about 5x slower. Process-wide switch, read at every forward.
_PRECISION = "bf16"


def set_precision(mode: str) -> str:
    """Select the arithmetic of every climb_b200 forward / backward issued from now on; returns the previous mode."""
    global _PRECISION
    if mode not in ("bf16", "bf16x3"):
        raise ValueError("precision must be 'bf16' (throughput) or 'bf16x3' (parity gate)")
    old, _PRECISION = _PRECISION, mode
    return old


def get_precision() -> str:
    return _PRECISION


class precision:
    """with ops.precision('bf16x3'): ..."""

    def __init__(self, mode: str):
        self.mode = mode

    def __enter__(self):
        self.old = set_precision(self.mode)
        return self

    def __exit__(self, *exc):
        set_precision(self.old)
        return False


def _bf16_padded(x: torch.Tensor, mult: int = 8) -> torch.Tensor:
    """bf16 copy of a 2-D fp32 tensor with the row stride padded to a multiple of 8 elements
    (TMA needs 16-byte row pitch); returns a [rows, cols] view of the padded buffer."""
    rows, cols = x.shape
    ld = (cols + mult - 1) // mult * mult
    buf = torch.zeros(rows, ld, dtype=torch.bfloat16, device=x.device)
    buf[:, :cols].copy_(x)
    return buf[:, :cols]


def _w_bf16(w: torch.Tensor) -> torch.Tensor:
    out = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
    if w.numel() % 8 == 0 and w.is_contiguous() and w.data_ptr() % 16 == 0:
        _lib.cast_f32_bf16(w.detach(), out)
    else:
        out.copy_(w.detach())
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM; backward = dgrad (W read in place, MN-major) + wgrad (both
    operands in place) + column-sum bias gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        xb = _bf16_padded(x2.detach().float())
        wb = _w_bf16(weight)
        N = weight.shape[0]
        out = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
        _lib.gemm(xb, wb, out, bias=bias.detach() if bias is not None else None)
        ctx.save_for_backward(xb, wb)
        ctx.has_bias = bias is not None
        ctx.lead = lead
        return out.view(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        N, K = wb.shape
        dyb = _bf16_padded(dy.reshape(-1, N).float())
        M = dyb.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
            _lib.gemm(dyb, wb, dx, b_mn_major=True, M=M, N=K, K=N)
            dx = dx.view(*ctx.lead, K)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(N, K, dtype=torch.float32, device=dy.device)
            _lib.gemm(dyb, xb, dw, a_mn_major=True, b_mn_major=True, accumulate=True, M=N, N=K, K=M)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, dtype=torch.float32, device=dy.device)
            _lib.colsum(dyb, db)
        return dx, dw, db


def _split_padded(x: torch.Tensor, mult: int = 8):
    """(hi, lo) bf16 views [rows, cols] of a 2-D fp32 tensor, row stride padded to a multiple of 8 elements."""
    rows, cols = x.shape
    ld = (cols + mult - 1) // mult * mult
    src = x.detach().float()
    if ld != cols:
        padded = torch.zeros(rows, ld, dtype=torch.float32, device=x.device)
        padded[:, :cols].copy_(src)
        src = padded
    src = src.contiguous()
    hi = torch.empty(rows, ld, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(rows, ld, dtype=torch.bfloat16, device=x.device)
    _lib.split_f32_bf16x2(src, hi, lo)
    return hi[:, :cols], lo[:, :cols]


def _gemm3(a, b, out, **kw):
    """out += Ahi Bhi^T + Alo Bhi^T + Ahi Blo^T: the split-operand contraction of the bf16x3 mode."""
    (a_hi, a_lo), (b_hi, b_lo) = a, b
    for x, y in ((a_hi, b_hi), (a_lo, b_hi), (a_hi, b_lo)):
        _lib.gemm(x, y, out, accumulate=True, **kw)
    return out


class _Linear3Fn(torch.autograd.Function):
    """_LinearFn in the bf16x3 precision mode (task heads, src/modeling/vilt.py:179-203)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        xs, ws = _split_padded(x2), _split_padded(weight)
        N = weight.shape[0]
        out = torch.zeros(x2.shape[0], N, dtype=torch.float32, device=x.device)
        if bias is not None:
            out += bias.detach()
        _gemm3(xs, ws, out)
        ctx.save_for_backward(*xs, *ws)
        ctx.has_bias, ctx.lead = bias is not None, lead
        return out.view(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        x_hi, x_lo, w_hi, w_lo = ctx.saved_tensors
        N, K = w_hi.shape
        dy2 = dy.reshape(-1, N).float().contiguous()
        dys = _split_padded(dy2)
        M = dy2.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.zeros(M, K, dtype=torch.float32, device=dy.device)
            _gemm3(dys, (w_hi, w_lo), dx, b_mn_major=True, M=M, N=K, K=N)
            dx = dx.view(*ctx.lead, K)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(N, K, dtype=torch.float32, device=dy.device)
            _gemm3(dys, (x_hi, x_lo), dw, a_mn_major=True, b_mn_major=True, M=N, N=K, K=M)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, dtype=torch.float32, device=dy.device)
            _lib.colsum(dy2, db)
        return dx, dw, db


def linear(x, weight, bias=None):
    if _PRECISION == "bf16x3":
        return _Linear3Fn.apply(x, weight, bias)
    return _LinearFn.apply(x, weight, bias)


class _LayerNormGeluFn(torch.autograd.Function):
    """LayerNorm (+ optional erf-GELU) in one kernel each way: the LN -> GELU pair of the
    classification heads (src/modeling/vilt.py:192-193)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, gelu):
        lead = x.shape[:-1]
        d = x.shape[-1]
        x2 = x.reshape(-1, d).contiguous().float()
        act = _lib.EPI_GELU if gelu else _lib.EPI_NONE
        _, y, mean, rstd = _lib.layernorm_fwd(x2, gamma.detach(), beta.detach(), eps, out_bf16=False, out_f32=True, act=act)
        ctx.save_for_backward(x2, gamma.detach(), beta.detach(), mean, rstd)
        ctx.act, ctx.lead = act, lead
        return y.view(*lead, d)

    @staticmethod
    def backward(ctx, dy):
        x2, gamma, beta, mean, rstd = ctx.saved_tensors
        d = x2.shape[-1]
        dy2 = dy.reshape(-1, d).contiguous().float()
        dx = torch.empty_like(x2)
        dg = torch.zeros(d, dtype=torch.float32, device=dy.device)
        db = torch.zeros(d, dtype=torch.float32, device=dy.device)
        _lib.layernorm_bwd(dy2, x2, gamma, beta, mean, rstd, dx_f32=dx, dgamma=dg, dbeta=db, act=ctx.act)
        return dx.view(*ctx.lead, d), dg, db, None, None


def layernorm_gelu(x, gamma, beta, eps=1e-5, gelu=True):
    return _LayerNormGeluFn.apply(x, gamma, beta, eps, gelu)


class _BCELossFn(torch.autograd.Function):
    """BCEWithLogitsLoss(mean) * scale with the gradient produced in the same pass (train_vqa.py:95,157)."""

    @staticmethod
    def forward(ctx, logits, target, scale):
        lg = logits.contiguous().float()
        tg = target.contiguous().float()
        rows, cols = lg.shape
        row_loss = torch.empty(rows, dtype=torch.float32, device=lg.device)
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        dlogits = torch.empty_like(lg)
        _lib.check(_lib.climb_bce_logits_loss(_lib.ptr(lg), cols, _lib.ptr(tg), rows, cols, float(scale), 1.0,
                                              _lib.ptr(row_loss), _lib.ptr(loss), _lib.ptr(dlogits), cols,
                                              _lib.stream()))
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None, None


class _CELossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        lg = logits.contiguous().float()
        tg = target.contiguous().to(torch.int64)
        rows, cols = lg.shape
        row_loss = torch.empty(rows, dtype=torch.float32, device=lg.device)
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        dlogits = torch.empty_like(lg)
        _lib.check(_lib.climb_cross_entropy_loss(_lib.ptr(lg), cols, _lib.ptr(tg), rows, cols, 1.0, _lib.ptr(row_loss),
                                                 _lib.ptr(loss), _lib.ptr(dlogits), cols, _lib.stream()))
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None


def vqa_loss(logits, target):
    """loss_criterion(logits, target) * target.shape[1] of train_vqa.py:157."""
    return _BCELossFn.apply(logits, target, float(target.shape[1]))


def cross_entropy_loss(logits, target):
    return _CELossFn.apply(logits, target)


class _EWCPenaltyFn(torch.autograd.Function):
    """lambda * sum F (theta - theta*)^2 over a flat arena; its backward adds 2 lambda F (theta - theta*)
    straight into the gradient arena in one streaming pass (ewc.py:75-87)."""

    @staticmethod
    def forward(ctx, anchor, arena, theta_star, fisher, lam, trainable, fisher_bwd=None):
        n = arena.size
        partials = torch.empty(2048, dtype=torch.float32, device=arena.theta.device)
        loss = torch.empty((), dtype=torch.float32, device=arena.theta.device)
        _lib.check(_lib.climb_ewc_penalty(_lib.ptr(arena.theta), _lib.ptr(theta_star), _lib.ptr(fisher), n, float(lam),
                                          _lib.ptr(partials), 2048, _lib.ptr(loss), None, 0.0, None, _lib.stream()))
        ctx.arena, ctx.theta_star, ctx.lam, ctx.trainable, ctx.partials = arena, theta_star, lam, trainable, partials
        ctx.fisher = fisher if fisher_bwd is None else fisher_bwd       # gradient weights: F restricted to the trainable slices
        return loss

    @staticmethod
    def backward(ctx, dloss):
        arena = ctx.arena
        if not ctx.trainable:           # no Fisher-tracked parameter trains (frozen base + adapters): nothing to add
            return None, None, None, None, None, None, None
        arena.prepare_grads(ctx.trainable)
        scratch = torch.empty((), dtype=torch.float32, device=arena.theta.device)
        # the upstream gradient (1 in the trainers' (loss + ewc_loss).backward()) is read on the device
        g = dloss.detach().to(torch.float32).contiguous()
        _lib.check(_lib.climb_ewc_penalty(_lib.ptr(arena.theta), _lib.ptr(ctx.theta_star), _lib.ptr(ctx.fisher), arena.size,
                                          float(ctx.lam), _lib.ptr(ctx.partials), 2048, _lib.ptr(scratch),
                                          _lib.ptr(arena.grad), 1.0, _lib.ptr(g), _lib.stream()))
        arena.publish_grads(ctx.trainable)
        return None, None, None, None, None, None, None


def ewc_penalty(arena, theta_star: torch.Tensor, fisher: torch.Tensor, lam: float, trainable, fisher_bwd=None):
    """fisher_bwd: F with the slices of frozen parameters zeroed -- the loss counts every Fisher-tracked parameter (as the
    reference's sum does), the gradient only reaches the trainable ones (autograd gives frozen leaves no .grad)."""
    anchor = torch.zeros((), device=arena.theta.device, requires_grad=True)
    return _EWCPenaltyFn.apply(anchor, arena, theta_star, fisher, lam, trainable, fisher_bwd)
