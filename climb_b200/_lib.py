"""ctypes binding of libclimb_b200.so (the C ABI declared in include/climb_b200.h).

The shared library is the product: there is no PyTorch / CPU fallback behind these calls. If the
library has not been built, importing this module raises immediately with the build command.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclimb_b200.so")

BF16, F32 = 0, 1
EPI_NONE, EPI_GELU, EPI_DGELU, EPI_SWISH, EPI_DSWISH, EPI_RELU, EPI_DRELU, EPI_TANH, EPI_GELU_SAVE_GRAD, EPI_MUL_AUX = range(10)


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
        ("colsum", c_void_p),
        ("alpha", c_float),
        ("accumulate", c_int),
        ("split_k", c_int),
        ("block_n", c_int),
        ("independent", c_int),
        ("colsum_a", c_void_p),
    ]


def _sig(name, argtypes, restype=c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


_P = c_void_p
climb_last_error = _sig("climb_last_error", [], c_char_p)
climb_version = _sig("climb_version", [])
climb_launch_count = _sig("climb_launch_count", [], ctypes.c_uint64)
climb_gemm_pair_mode = _sig("climb_gemm_pair_mode", [ctypes.c_int])
climb_set_sm_reserve = _sig("climb_set_sm_reserve", [ctypes.c_int])
climb_wordpiece_create = _sig("climb_wordpiece_create", [c_char_p, ctypes.c_int64, c_int, c_int], c_void_p)
climb_wordpiece_destroy = _sig("climb_wordpiece_destroy", [c_void_p], None)
climb_wordpiece_encode = _sig("climb_wordpiece_encode", [c_void_p, c_char_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                                         ctypes.POINTER(c_int), c_int])
climb_profile_begin = _sig("climb_profile_begin", [])
climb_profile_end = _sig("climb_profile_end", [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int64), c_int])
climb_gemm_bf16 = _sig("climb_gemm_bf16", [POINTER(GemmDesc), _P])
climb_adapter_fused = _sig("climb_adapter_fused", [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P])
climb_attention_fwd = _sig("climb_attention_fwd", [_P, _P, _P, _P, c_int, c_int, c_int, c_float, _P])
climb_attention_fwd_dropout = _sig("climb_attention_fwd_dropout",
                                   [_P, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, ctypes.c_uint64, _P])
climb_attention_bwd = _sig(
    "climb_attention_bwd", [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P])
climb_layernorm_fwd = _sig(
    "climb_layernorm_fwd", [_P, c_int64, _P, _P, c_float, _P, _P, _P, _P, c_int, c_int, c_int, _P])
climb_layernorm_bwd = _sig(
    "climb_layernorm_bwd",
    [_P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P])


climb_layernorm_bwd_colsum = _sig(
    "climb_layernorm_bwd_colsum",
    [_P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P])
climb_cast_f32_bf16 = _sig("climb_cast_f32_bf16", [_P, _P, c_int64, _P])
climb_split_f32_bf16x2 = _sig("climb_split_f32_bf16x2", [_P, _P, _P, c_int64, _P])
climb_error_flags = _sig("climb_error_flags", [], ctypes.c_uint32)
climb_colsum = _sig("climb_colsum", [_P, c_int, c_int64, c_int, c_int, _P, _P])
climb_bce_logits_loss = _sig(
    "climb_bce_logits_loss", [_P, c_int64, _P, c_int, c_int, c_float, c_float, _P, _P, _P, c_int64, _P])
climb_cross_entropy_loss = _sig(
    "climb_cross_entropy_loss", [_P, c_int64, _P, c_int, c_int, c_float, _P, _P, _P, c_int64, _P])
climb_ewc_penalty = _sig(
    "climb_ewc_penalty", [_P, _P, _P, c_int64, c_float, _P, c_int, _P, _P, c_float, _P, _P])
climb_fisher_accumulate = _sig("climb_fisher_accumulate", [_P, _P, c_int64, _P])
climb_scale_inplace = _sig("climb_scale_inplace", [_P, c_int64, c_float, _P])
climb_adamw_step = _sig(
    "climb_adamw_step", [_P, _P, _P, _P, _P, _P, c_int, POINTER(c_float), POINTER(c_float), c_int, c_float, c_float,
                         c_float, c_int, _P])


class ViltDimsC(Structure):
    _fields_ = [("hidden", c_int), ("layers", c_int), ("heads", c_int), ("ffn", c_int),
                ("patch", c_int), ("channels", c_int), ("pos_grid", c_int), ("n_modality", c_int),
                ("ln_eps", c_float), ("vocab_size", c_int), ("type_vocab_size", c_int), ("precision", c_int),
                ("hidden_dropout", c_float), ("attn_dropout", c_float)]


PREC_BF16, PREC_BF16X3 = 0, 1
ERR_TOKEN_ID, ERR_TOKEN_TYPE, ERR_MODALITY = 1, 2, 4


LAYER_FIELDS = ["qkv_w", "qkv_b", "o_w", "o_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                "ln1_w", "ln1_b", "ln2_w", "ln2_b",
                "mh_down_w", "mh_down_b", "mh_up_w", "mh_up_b",
                "out_down_w", "out_down_b", "out_up_w", "out_up_b"]


class ViltLayerC(Structure):
    _fields_ = [(n, c_int64) for n in LAYER_FIELDS] + [("flags", c_int32), ("pad_", c_int32)]


PARAM_FIELDS = ["cls_token", "pos_emb", "word_emb", "text_pos_emb", "text_type_emb", "text_ln_w",
                "text_ln_b", "patch_w", "patch_b", "mod_emb", "final_ln_w", "final_ln_b", "pooler_w",
                "pooler_b"]


class ViltParamsC(Structure):
    _fields_ = [(n, c_int64) for n in PARAM_FIELDS] + [
        ("layer", POINTER(ViltLayerC)), ("adapter_r", c_int), ("adapter_act", c_int),
        ("embed_flags", c_int32), ("tail_flags", c_int32), ("shadow_lo", c_void_p)]


class ViltBatchC(Structure):
    _fields_ = [("B", c_int), ("T", c_int), ("H", c_int), ("W", c_int),
                ("input_ids", c_void_p), ("inputs_embeds", c_void_p), ("token_type_ids", c_void_p),
                ("attention_mask", c_void_p), ("pixel_values", c_void_p), ("image_type_idx", c_void_p),
                ("image_type_idx_scalar", c_int), ("patch_geom", c_void_p), ("n_patch_slots", c_int),
                ("training", c_int), ("dropout_seed", ctypes.c_uint64), ("image_repeat", c_int), ("patch_select", c_void_p)]


class AdamWChunkC(Structure):
    _fields_ = [("start", c_int64), ("length", c_int32), ("group", c_int32)]


TRAIN_BASE, TRAIN_ADAPTER = 1, 2
BWD_TAIL, BWD_EMBED = 1, 2

climb_vilt_forward_workspace_bytes = _sig(
    "climb_vilt_forward_workspace_bytes", [POINTER(ViltDimsC), POINTER(ViltParamsC), POINTER(ViltBatchC), c_int],
    c_int64)
climb_vilt_backward_scratch_bytes = _sig(
    "climb_vilt_backward_scratch_bytes", [POINTER(ViltDimsC), POINTER(ViltParamsC), POINTER(ViltBatchC)], c_int64)
climb_vilt_forward = _sig(
    "climb_vilt_forward", [POINTER(ViltDimsC), POINTER(ViltParamsC), POINTER(ViltBatchC), _P, _P, _P, c_int64,
                           c_int, _P, _P])
climb_vilt_backward = _sig(
    "climb_vilt_backward", [POINTER(ViltDimsC), POINTER(ViltParamsC), POINTER(ViltBatchC), _P, _P, _P, c_int64,
                            _P, c_int64, _P, _P, c_int, c_int, c_int, _P])


# ---- frozen BERT text encoder (ViLT-BERT) -----------------------------------------------------------
class BertDimsC(Structure):
    _fields_ = [("hidden", c_int), ("layers", c_int), ("heads", c_int), ("ffn", c_int), ("ln_eps", c_float)]


BERT_LAYER_FIELDS = ["qkv_w", "qkv_b", "o_w", "o_b", "attn_ln_w", "attn_ln_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                     "out_ln_w", "out_ln_b"]


class BertLayerC(Structure):
    _fields_ = [(n, c_int64) for n in BERT_LAYER_FIELDS]


class BertParamsC(Structure):
    _fields_ = [(n, c_int64) for n in ("word_emb", "pos_emb", "type_emb", "emb_ln_w", "emb_ln_b")] + [
        ("layer", POINTER(BertLayerC))]


class BertBatchC(Structure):
    _fields_ = [("B", c_int), ("T", c_int), ("input_ids", c_void_p), ("token_type_ids", c_void_p),
                ("attention_mask", c_void_p)]


class ImageDescC(Structure):
    """climb_image_desc (include/climb_b200.h)."""
    _fields_ = [("src_off", c_int64), ("tmp_off", c_int64), ("in_h", ctypes.c_int32), ("in_w", ctypes.c_int32),
                ("out_h", ctypes.c_int32), ("out_w", ctypes.c_int32), ("ksize_h", ctypes.c_int32), ("ksize_v", ctypes.c_int32),
                ("bounds_h_off", c_int64), ("coef_h_off", c_int64), ("bounds_v_off", c_int64), ("coef_v_off", c_int64)]


climb_image_preprocess = _sig("climb_image_preprocess", [_P, _P, _P, _P, c_int, c_int64, _P, _P, c_int, c_int,
                                                         POINTER(c_float), POINTER(c_float), _P])

climb_bert_forward_workspace_bytes = _sig("climb_bert_forward_workspace_bytes", [POINTER(BertDimsC), POINTER(BertBatchC)],
                                          c_int64)
climb_bert_forward = _sig("climb_bert_forward", [POINTER(BertDimsC), POINTER(BertParamsC), POINTER(BertBatchC), _P, _P, _P,
                                                 c_int64, c_float, c_float, ctypes.c_uint64, _P, _P])
climb_dropout_add = _sig("climb_dropout_add", [_P, _P, _P, c_int64, c_float, ctypes.c_uint64, _P])
climb_dropout_site_seed = _sig("climb_dropout_site_seed", [ctypes.c_uint64, c_int, c_int], ctypes.c_uint64)
climb_dropout_keep_mask = _sig("climb_dropout_keep_mask", [_P, c_int64, c_float, ctypes.c_uint64, _P])
climb_attention_dropout_keep_mask = _sig("climb_attention_dropout_keep_mask", [_P, c_int, c_int, c_int, c_float, ctypes.c_uint64, _P])


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
         bias=None, residual=None, epilogue=EPI_NONE, aux=None, c2=None, colsum=None, alpha=1.0, accumulate=False,
         split_k=0, block_n=0, M=None, N=None, K=None, colsum_a=None) -> torch.Tensor:
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
    d.colsum = ptr(colsum)
    d.alpha = alpha
    d.accumulate = int(accumulate)
    d.split_k = split_k
    d.block_n = block_n
    d.colsum_a = ptr(colsum_a)
    if colsum_a is not None:
        assert colsum_a.dtype == torch.float32 and colsum_a.numel() >= M
    check(climb_gemm_bf16(ctypes.byref(d), stream()))
    return out


def attention_fwd(qkv, key_bias, B, L, H, scale, p_drop=0.0, seed=0):
    ctx = torch.empty(B, L, H * 64, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, H, L, dtype=torch.float32, device=qkv.device)
    if p_drop > 0.0:
        check(climb_attention_fwd_dropout(ptr(qkv), ptr(key_bias), ptr(ctx), ptr(lse), B, L, H, scale, p_drop, seed, stream()))
        return ctx, lse
    check(climb_attention_fwd(ptr(qkv), ptr(key_bias), ptr(ctx), ptr(lse), B, L, H, scale, stream()))
    return ctx, lse


def attention_bwd(qkv, key_bias, ctx, dctx, lse, B, L, H, scale, colsum=None):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    check(climb_attention_bwd(ptr(qkv), ptr(key_bias), ptr(ctx), ptr(dctx), ptr(lse), ptr(delta),
                              ptr(dqkv), ptr(colsum), B, L, H, scale, stream()))
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
                  dx_bf16=None, dgamma=None, dbeta=None, act=EPI_NONE, dx_colsum=None):
    d = gamma.numel()
    if rows is None:
        rows = mean.numel()
    if ldx is None:
        ldx = d
    dy_f32 = dy if dy.dtype == torch.float32 else None
    dy_b16 = dy if dy.dtype == torch.bfloat16 else None
    if dx_colsum is not None:
        check(climb_layernorm_bwd_colsum(ptr(dy_f32), ptr(dy_b16), ptr(x), ldx, ptr(gamma), ptr(beta), ptr(mean),
                                         ptr(rstd), ptr(dres), ptr(dx_f32), ptr(dx_bf16), ptr(dgamma), ptr(dbeta),
                                         ptr(dx_colsum), rows, d, act, stream()))
        return
    check(climb_layernorm_bwd(ptr(dy_f32), ptr(dy_b16), ptr(x), ldx, ptr(gamma), ptr(beta), ptr(mean),
                              ptr(rstd), ptr(dres), ptr(dx_f32), ptr(dx_bf16), ptr(dgamma), ptr(dbeta),
                              rows, d, act, stream()))


def cast_f32_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    check(climb_cast_f32_bf16(ptr(src), ptr(dst), src.numel(), stream()))


def split_f32_bf16x2(src: torch.Tensor, hi, lo) -> None:
    """hi = bf16(src), lo = bf16(src - hi): the split operands of the bf16x3 precision mode (either may be None)."""
    assert src.dtype == torch.float32 and src.is_contiguous()
    check(climb_split_f32_bf16x2(ptr(src), ptr(hi), ptr(lo), src.numel(), stream()))


def raise_device_errors() -> None:
    """Sticky device-side error flags -> Python exceptions (the reference's nn.Embedding raises IndexError for an
    out-of-range id; our kernels clamp the id, raise a flag and keep going -- the flag of a kernel becomes visible once it
    has run, so this reports at the next forward / at an explicit call)."""
    bits = climb_error_flags()
    if bits:
        what = [n for b, n in ((ERR_TOKEN_ID, "input_ids outside the word-embedding table"),
                               (ERR_TOKEN_TYPE, "token_type_ids outside the token-type table"),
                               (ERR_MODALITY, "image_token_type_idx outside the modality-type table")) if bits & b]
        raise IndexError("index out of range in an embedding lookup of an earlier climb_b200 forward: " + "; ".join(what))


def colsum(src: torch.Tensor, out: torch.Tensor, rows=None, cols=None) -> None:
    """out[c] += sum_r src[r, c] (src bf16 or fp32, row stride src.stride(0))."""
    assert src.dim() == 2 and src.stride(1) == 1 and out.dtype == torch.float32
    rows = src.shape[0] if rows is None else rows
    cols = src.shape[1] if cols is None else cols
    dt = {torch.bfloat16: BF16, torch.float32: F32}[src.dtype]
    check(climb_colsum(ptr(src), dt, src.stride(0), rows, cols, ptr(out), stream()))


def profile_end():
    """-> {category: (ms, work, launches)} for gemm / attn_fwd / attn_bwd."""
    ms = (ctypes.c_double * 4)()
    work = (ctypes.c_double * 4)()
    n = (c_int64 * 4)()
    check(climb_profile_end(ms, work, n, 4))
    names = ["gemm", "attn_fwd", "attn_bwd", "other"]
    return {names[i]: (ms[i], work[i], n[i]) for i in range(4)}
