"""ctypes binding of libclimb_b200.so (the C ABI declared in include/climb_b200.h).

The shared library is the product: there is no PyTorch / CPU fallback behind these calls. If the
library has not been built, importing this module raises immediately with the build command.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclimb_b200.so")

BF16, F32 = 0, 1
EPI_NONE, EPI_GELU, EPI_DGELU, EPI_SWISH, EPI_DSWISH, EPI_RELU, EPI_DRELU, EPI_TANH = range(8)


class ClimbError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C climb_b200/csrc`). "
            "climb_b200 has no CPU fallback."
        )
    return ctypes.CDLL(LIB_PATH)


lib = _load()


class GemmDesc(Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_void_p), ("lda", c_int64), ("a_mn_major", c_int),
        ("B", c_void_p), ("ldb", c_int64), ("b_mn_major", c_int),
        ("C", c_void_p), ("ldc", c_int64), ("c_dtype", c_int),
        ("bias", c_void_p),
        ("residual", c_void_p), ("ldr", c_int64),
        ("epilogue", c_int),
        ("aux", c_void_p), ("ldaux", c_int64),
        ("c2", c_void_p), ("ldc2", c_int64),
        ("alpha", c_float),
        ("accumulate", c_int),
        ("split_k", c_int),
        ("block_n", c_int),
    ]


def _sig(name, argtypes, restype=c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


_P = c_void_p
climb_last_error = _sig("climb_last_error", [], c_char_p)
climb_version = _sig("climb_version", [])
climb_gemm_bf16 = _sig("climb_gemm_bf16", [POINTER(GemmDesc), _P])
climb_attention_fwd = _sig("climb_attention_fwd", [_P, _P, _P, _P, c_int, c_int, c_int, c_float, _P])
climb_attention_bwd = _sig(
    "climb_attention_bwd", [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P])
climb_layernorm_fwd = _sig(
    "climb_layernorm_fwd", [_P, c_int64, _P, _P, c_float, _P, _P, _P, _P, c_int, c_int, c_int, _P])
climb_layernorm_bwd = _sig(
    "climb_layernorm_bwd",
    [_P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P])


def check(rc: int) -> None:
    if rc != 0:
        msg = climb_last_error()
        raise ClimbError(f"libclimb_b200 call failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t: torch.Tensor | None):
    """Device pointer of a tensor (None -> NULL). The tensor must be a CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ClimbError("climb_b200 kernels take CUDA tensors only (no CPU fallback)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# -------------------------------------------------------------------------------------------------
# Thin tensor-level wrappers (no autograd here; see climb_b200.ops)
# -------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, a_mn_major=False, b_mn_major=False,
         bias=None, residual=None, epilogue=EPI_NONE, aux=None, c2=None, alpha=1.0, accumulate=False,
         split_k=0, block_n=0, M=None, N=None, K=None) -> torch.Tensor:
    """out[M,N] = epi(alpha * A B^T + bias) + residual. A, B bf16 2-D (possibly row-strided views)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    if M is None:
        M = a.shape[1] if a_mn_major else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn_major else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn_major else b.shape[0]
    kb = b.shape[0] if b_mn_major else b.shape[1]
    assert kb == K, f"contraction mismatch: A has K={K}, B has K={kb}"
    assert out.shape[0] >= M and out.shape[1] >= N
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn_major = ptr(a), a.stride(0), int(a_mn_major)
    d.B, d.ldb, d.b_mn_major = ptr(b), b.stride(0), int(b_mn_major)
    d.C, d.ldc = ptr(out), out.stride(0)
    d.c_dtype = {torch.bfloat16: BF16, torch.float32: F32}[out.dtype]
    d.bias = ptr(bias)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    d.residual = ptr(residual)
    d.ldr = residual.stride(0) if residual is not None else 0
    if residual is not None:
        assert residual.dtype == torch.float32
    d.epilogue = epilogue
    d.aux = ptr(aux)
    d.ldaux = aux.stride(0) if aux is not None else 0
    if aux is not None:
        assert aux.dtype == torch.bfloat16
    d.c2 = ptr(c2)
    d.ldc2 = c2.stride(0) if c2 is not None else 0
    if c2 is not None:
        assert c2.dtype == torch.bfloat16
    d.alpha = alpha
    d.accumulate = int(accumulate)
    d.split_k = split_k
    d.block_n = block_n
    check(climb_gemm_bf16(ctypes.byref(d), stream()))
    return out


def attention_fwd(qkv, key_bias, B, L, H, scale):
    ctx = torch.empty(B, L, H * 64, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, H, L, dtype=torch.float32, device=qkv.device)
    check(climb_attention_fwd(ptr(qkv), ptr(key_bias), ptr(ctx), ptr(lse), B, L, H, scale, stream()))
    return ctx, lse


def attention_bwd(qkv, key_bias, ctx, dctx, lse, B, L, H, scale):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    check(climb_attention_bwd(ptr(qkv), ptr(key_bias), ptr(ctx), ptr(dctx), ptr(lse), ptr(delta),
                              ptr(dqkv), B, L, H, scale, stream()))
    return dqkv


def layernorm_fwd(x, gamma, beta, eps, *, rows=None, ldx=None, out_bf16=True, out_f32=False,
                  act=EPI_NONE):
    d = gamma.numel()
    if rows is None:
        rows = x.numel() // d
    if ldx is None:
        ldx = d
    yb = torch.empty(rows, d, dtype=torch.bfloat16, device=x.device) if out_bf16 else None
    yf = torch.empty(rows, d, dtype=torch.float32, device=x.device) if out_f32 else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    check(climb_layernorm_fwd(ptr(x), ldx, ptr(gamma), ptr(beta), eps, ptr(yb), ptr(yf), ptr(mean),
                              ptr(rstd), rows, d, act, stream()))
    return yb, yf, mean, rstd


def layernorm_bwd(dy, x, gamma, beta, mean, rstd, *, rows=None, ldx=None, dres=None, dx_f32=None,
                  dx_bf16=None, dgamma=None, dbeta=None, act=EPI_NONE):
    d = gamma.numel()
    if rows is None:
        rows = mean.numel()
    if ldx is None:
        ldx = d
    dy_f32 = dy if dy.dtype == torch.float32 else None
    dy_b16 = dy if dy.dtype == torch.bfloat16 else None
    check(climb_layernorm_bwd(ptr(dy_f32), ptr(dy_b16), ptr(x), ldx, ptr(gamma), ptr(beta), ptr(mean),
                              ptr(rstd), ptr(dres), ptr(dx_f32), ptr(dx_bf16), ptr(dgamma), ptr(dbeta),
                              rows, d, act, stream()))
