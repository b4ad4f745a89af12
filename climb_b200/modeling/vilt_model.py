"""B200ViltModel: the drop-in for adapter-transformers' ViltModel on CLiMB's hot path.

Same parameter tree (names, shapes, registration order) as
`transformers.models.vilt.modeling_vilt.ViltModel` (modeling_vilt.py:744-884) so that HF
`dandelin/vilt-b32-*` weights and CLiMB checkpoints load unchanged, same
forward signature and `pooler_output`, same adapter API (`add_adapter / train_adapter /
set_active_adapters / active_adapters`, adapters/model_mixin.py:156-274) -- but forward and backward
are ONE call each into libclimb_b200.so (include/climb_b200.h: climb_vilt_forward /
climb_vilt_backward). The nn.Linear / nn.LayerNorm / nn.Embedding / nn.Conv2d children below are
parameter containers only: their own forward() is never called and there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ..arena import ParamArena


@dataclass
class B200ViltConfig:
    """Fields of ViltConfig (configuration_vilt.py:101-124) that the hot path reads."""
    vocab_size: int = 30522
    type_vocab_size: int = 2
    modality_type_vocab_size: int = 2
    max_position_embeddings: int = 40
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    hidden_dropout_prob: float = 0.0
    attention_probs_dropout_prob: float = 0.0
    initializer_range: float = 0.02
    layer_norm_eps: float = 1e-12
    image_size: int = 384
    patch_size: int = 32
    num_channels: int = 3
    qkv_bias: bool = True
    max_image_length: int = -1

    @classmethod
    def from_hf(cls, cfg) -> "B200ViltConfig":
        """Accepts a transformers ViltConfig (or any object / dict with the same attributes)."""
        get = (lambda k, dflt: cfg.get(k, dflt)) if isinstance(cfg, dict) else (lambda k, dflt: getattr(cfg, k, dflt))
        return cls(**{f: get(f, getattr(cls, f)) for f in cls.__dataclass_fields__})


@dataclass
class AdapterSpec:
    """What CLiMB's AdapterHandler configures (src/cl_algorithms/adapters.py:27-50) reduced to the
    bottleneck the ViLT mixins actually execute (adapters/mixins/vilt.py:23-125, modeling.py:120-201)."""
    reduction_factor: float = 16
    non_linearity: str = "swish"
    mh_adapter: bool = True
    output_adapter: bool = True
    scaling: float = 1.0

    @classmethod
    def from_config(cls, config) -> "AdapterSpec":
        if isinstance(config, AdapterSpec):
            return config
        if isinstance(config, str):
            presets = {"houlsby": cls(16, "swish", True, True), "pfeiffer": cls(16, "relu", False, True)}
            if config not in presets:
                raise ValueError(f"unsupported adapter config '{config}' (supported: {sorted(presets)})")
            return presets[config]
        get = (lambda k, d=None: config.get(k, d)) if hasattr(config, "get") else (lambda k, d=None: getattr(config, k, d))
        spec = cls(reduction_factor=get("reduction_factor", 16), non_linearity=str(get("non_linearity", "swish")).lower(),
                   mh_adapter=bool(get("mh_adapter", True)), output_adapter=bool(get("output_adapter", True)),
                   scaling=get("scaling", 1.0))
        unsupported = []
        if get("phm_layer", False):
            unsupported.append("phm_layer (compacter)")
        if get("ln_before", False) or get("ln_after", False):
            unsupported.append("adapter LayerNorm")
        if get("is_parallel", False):
            unsupported.append("parallel adapters (degenerate on ViLT: SURVEY.md 3.5)")
        if get("inv_adapter", None):
            unsupported.append("invertible adapters")
        if not isinstance(spec.scaling, (int, float)) or float(spec.scaling) != 1.0:
            unsupported.append("scaling != 1.0")
        if get("residual_before_ln", True) is not True or get("adapter_residual_before_ln", False):
            unsupported.append("non-default residual placement")
        if spec.non_linearity not in ("swish", "silu", "relu"):
            unsupported.append(f"non_linearity={spec.non_linearity}")
        if unsupported:
            raise NotImplementedError("climb_b200 adapters cover the Houlsby / Pfeiffer bottlenecks CLiMB ships "
                                      f"scripts for; unsupported options: {', '.join(unsupported)}")
        if spec.non_linearity == "silu":
            spec.non_linearity = "swish"
        return spec


class _Adapter(nn.Module):
    """Parameter container named like adapters/modeling.py:31-125 (adapter_down.0 / adapter_up)."""

    def __init__(self, d: int, r: int, init_range: float = 0.02):
        super().__init__()
        self.adapter_down = nn.Sequential(nn.Linear(d, r), nn.Identity())
        self.adapter_up = nn.Linear(r, d)
        for lin in (self.adapter_down[0], self.adapter_up):      # init_bert_weights, modeling.py:204-214
            nn.init.normal_(lin.weight, mean=0.0, std=init_range)
            nn.init.zeros_(lin.bias)


class _SelfAttention(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.query = nn.Linear(c.hidden_size, c.hidden_size, bias=c.qkv_bias)
        self.key = nn.Linear(c.hidden_size, c.hidden_size, bias=c.qkv_bias)
        self.value = nn.Linear(c.hidden_size, c.hidden_size, bias=c.qkv_bias)


class _SelfOutput(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.adapters = nn.ModuleDict()


class _Attention(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.attention = _SelfAttention(c)
        self.output = _SelfOutput(c)


class _Intermediate(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)


class _Output(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.dense = nn.Linear(c.intermediate_size, c.hidden_size)
        self.adapters = nn.ModuleDict()


class _Layer(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.attention = _Attention(c)
        self.intermediate = _Intermediate(c)
        self.output = _Output(c)
        self.layernorm_before = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.layernorm_after = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(c) for _ in range(c.num_hidden_layers)])


class _TextEmbeddings(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(c.max_position_embeddings).expand((1, -1)))


class _PatchEmbeddings(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.projection = nn.Conv2d(c.num_channels, c.hidden_size, kernel_size=c.patch_size, stride=c.patch_size)


class _Embeddings(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.text_embeddings = _TextEmbeddings(c)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, c.hidden_size))
        self.patch_embeddings = _PatchEmbeddings(c)
        n_patches = (c.image_size // c.patch_size) ** 2
        self.position_embeddings = nn.Parameter(torch.zeros(1, n_patches + 1, c.hidden_size))
        self.token_type_embeddings = nn.Embedding(c.modality_type_vocab_size, c.hidden_size)


class _Pooler(nn.Module):
    def __init__(self, c: B200ViltConfig):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)


@dataclass
class ViltOutput:
    """BaseModelOutputWithPooling, reduced to what the hot path produces: CLiMB reads only
    pooler_output (src/modeling/vilt.py:123-124)."""
    pooler_output: torch.Tensor
    last_hidden_state: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class _EncoderFn(torch.autograd.Function):
    """Autograd node of the whole encoder. Parameter gradients are ACCUMULATED by the engine straight
    into the flat gradient arena (whose slices are the parameters' .grad tensors); the node's only
    differentiable input is an anchor scalar that keeps it in the graph."""

    @staticmethod
    def forward(ctx, anchor, model, call):
        pooled = model._run_forward(call, save=True)
        ctx.model, ctx.call = model, call
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        ctx.model._run_backward(ctx.call, dpooled.contiguous().float())
        ctx.call = None
        return None, None, None


class _Call:
    """Everything one forward/backward pair shares: C structs, the tensors they point into, workspace."""
    __slots__ = ("dims", "params", "layers", "batch", "keep", "workspace", "ws_bytes", "arena_theta", "trainable",
                 "param_version")


class B200ViltModel(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        self.config = config if isinstance(config, B200ViltConfig) else B200ViltConfig.from_hf(config or {})
        c = self.config
        if c.hidden_size != c.num_attention_heads * 64 or c.hidden_size % 128:
            raise ValueError("climb_b200 kernels need head_dim 64 and hidden_size % 128 == 0 "
                             f"(got hidden={c.hidden_size}, heads={c.num_attention_heads})")
        if not (0.0 <= c.hidden_dropout_prob < 1.0 and 0.0 <= c.attention_probs_dropout_prob < 1.0):
            raise ValueError("dropout probabilities must lie in [0, 1)")
        if c.hidden_act != "gelu":
            raise NotImplementedError("only hidden_act='gelu' (erf) is implemented")
        self.embeddings = _Embeddings(c)
        self.encoder = _Encoder(c)
        self.layernorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.pooler = _Pooler(c)
        self.adapter_specs: Dict[str, AdapterSpec] = {}
        self._active_adapter: Optional[str] = None
        self._arena = ParamArena(self, "_arena_items")
        self._scratch: Optional[torch.Tensor] = None
        self.grad_sync = None          # set by climb_b200.distributed: callable(arena) after each backward
        self.apply(self._init_weights)

    # -- initialisation as modeling_vilt.py:597-611 ------------------------------------------------
    def _init_weights(self, m):
        std = self.config.initializer_range
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)

    # -- arena order: network order (embeddings, layer 0 .. layer L-1, tail) so that the backward pass
    #    fills the gradient arena back to front; inside a layer q,k,v weights (and biases) are adjacent
    #    so that one [3d, d] GEMM reads them ----------------------------------------------------------
    def _arena_items(self) -> List[Tuple[str, nn.Parameter]]:
        named = dict(self.named_parameters())
        order: List[str] = [n for n in named if n.startswith("embeddings.")]
        for i in range(len(self.encoder.layer)):
            pre = f"encoder.layer.{i}."
            a = pre + "attention.attention."
            qkv = [a + "query.weight", a + "key.weight", a + "value.weight"]
            if self.config.qkv_bias:
                qkv += [a + "query.bias", a + "key.bias", a + "value.bias"]
            order += qkv
            skip = set(qkv)
            order += [n for n in named if n.startswith(pre) and n not in skip]
        seen = set(order)
        order += [n for n in named if n not in seen]
        return [(n, named[n]) for n in order]

    # -- adapter API (adapters/model_mixin.py:156-274, as CLiMB calls it: src/modeling/vilt.py:356-367) --
    def add_adapter(self, adapter_name: str, config=None, overwrite_ok: bool = False, set_active: bool = False):
        spec = AdapterSpec.from_config(config if config is not None else "houlsby")
        if adapter_name in self.adapter_specs and not overwrite_ok:
            raise ValueError(f"Adapter '{adapter_name}' already exists.")
        d = self.config.hidden_size
        r = max(1, int(d // spec.reduction_factor))
        if r % 8:
            raise NotImplementedError(f"adapter bottleneck width {r} must be a multiple of 8 (TMA row alignment)")
        dev = next(self.parameters()).device
        for layer in self.encoder.layer:
            if spec.mh_adapter:
                layer.attention.output.adapters[adapter_name] = _Adapter(d, r, self.config.initializer_range).to(dev)
            if spec.output_adapter:
                layer.output.adapters[adapter_name] = _Adapter(d, r, self.config.initializer_range).to(dev)
        self.adapter_specs[adapter_name] = spec
        if set_active:
            self.set_active_adapters(adapter_name)

    def train_adapter(self, adapter_setup, train_embeddings: bool = False):
        """Freeze the base model, unfreeze the named adapter and activate it (model_mixin.py:156-173)."""
        name = self._single_adapter_name(adapter_setup)
        self.train()
        for n, p in self.named_parameters():
            p.requires_grad = (f".adapters.{name}." in n) or (train_embeddings and n.startswith("embeddings."))
        self.set_active_adapters(name)

    def set_active_adapters(self, adapter_setup):
        name = self._single_adapter_name(adapter_setup) if adapter_setup is not None else None
        if name is not None and name not in self.adapter_specs:
            raise ValueError(f"No adapter with name '{name}' found. Please make sure that all specified adapters are correctly loaded.")
        self._active_adapter = name

    @property
    def active_adapters(self):
        return self._active_adapter

    @staticmethod
    def _single_adapter_name(setup) -> str:
        if isinstance(setup, str):
            return setup
        if isinstance(setup, (list, tuple)) and len(setup) == 1 and isinstance(setup[0], str):
            return setup[0]
        children = getattr(setup, "children", None)          # an adapter-transformers Stack[...] block
        if children is not None and len(children) == 1 and isinstance(children[0], str):
            return children[0]
        raise NotImplementedError("climb_b200 runs one active adapter (Stack[name]), which is all CLiMB uses; "
                                  f"got {setup!r}")

    def freeze_model(self, freeze: bool = True):
        for p in self.parameters():
            p.requires_grad = not freeze

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, pixel_values=None, pixel_mask=None,
                head_mask=None, inputs_embeds=None, image_embeds=None, image_token_type_idx=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, image_repeat: int = 1,
                patch_draw_order=None):
        """ViltModel.forward (modeling_vilt.py:777-884). image_token_type_idx may be an int (as in the reference) or an
        int tensor [B] (batched NLVR2 passes). image_repeat = r > 1 (extension): pixel_values / pixel_mask hold B / r images,
        image i belongs to the r consecutive text rows i * r ... (VCR's four answer choices over one image,
        src/modeling/vilt.py:334-347): same result as pixel_values.repeat_interleave(r, 0), patch projection once per image.
        patch_draw_order (extension, only read when config.max_image_length > 0 drops patches): the order in which the
        sequences of this batched call would have drawn their random patch subsets in the reference's separate encoder
        passes (NLVR2: image 0 of every sample, then image 1; VCR: choice 0 of every sample, then choice 1 ...)."""
        if head_mask is not None or image_embeds is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("head_mask / image_embeds / output_attentions / output_hidden_states are outside "
                                      "CLiMB's hot path and not implemented in climb_b200")
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You have to specify exactly one of input_ids or inputs_embeds")
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        call = self._prepare_call(input_ids, attention_mask, token_type_ids, pixel_values, pixel_mask, inputs_embeds,
                                  image_token_type_idx, int(image_repeat), patch_draw_order)
        if torch.is_grad_enabled() and call.trainable:
            anchor = torch.zeros((), device=pixel_values.device, requires_grad=True)
            pooled = _EncoderFn.apply(anchor, self, call)
        else:
            pooled = self._run_forward(call, save=False)
        return ViltOutput(pooler_output=pooled)

    def _prepare_call(self, input_ids, attention_mask, token_type_ids, pixel_values, pixel_mask, inputs_embeds,
                      image_token_type_idx, image_repeat: int = 1, draw_order=None) -> _Call:
        c = self.config
        dev = pixel_values.device
        if dev.type != "cuda":
            raise _lib.ClimbError("climb_b200 runs on CUDA tensors only (no CPU fallback)")
        arena = self._arena
        arena.sync(dev)
        _lib.raise_device_errors()          # an out-of-range id of an EARLIER forward (clamped on the device) surfaces here
        Bi, C, H, W = pixel_values.shape
        text = input_ids if input_ids is not None else inputs_embeds
        T = text.shape[1]
        if image_repeat < 1:
            raise ValueError(f"image_repeat must be >= 1 (got {image_repeat})")
        from .. import ops as _ops
        if image_repeat > 1 and _ops.get_precision() == "bf16x3":
            # the precise (bf16x3) engine keeps one image per sequence
            pixel_values = pixel_values.repeat_interleave(image_repeat, dim=0)
            pixel_mask = None if pixel_mask is None else pixel_mask.repeat_interleave(image_repeat, dim=0)
            Bi, image_repeat = Bi * image_repeat, 1
        B = Bi * image_repeat
        if text.shape[0] != B:
            raise ValueError(f"{text.shape[0]} text rows for {Bi} images x image_repeat {image_repeat}")
        if C != c.num_channels or H % c.patch_size or W % c.patch_size:
            raise NotImplementedError(f"pixel_values {tuple(pixel_values.shape)}: the fixed-resolution path needs "
                                      f"{c.num_channels} channels and H, W multiples of {c.patch_size}")
        sel = None
        if isinstance(c.max_image_length, int) and 0 < c.max_image_length < (H // c.patch_size) * (W // c.patch_size):
            # the cap may bite (modeling_vilt.py:163-189). Every encoder call of the reference draws its own subset, so images
            # shared by several sequences (VCR) are expanded first
            if image_repeat > 1:
                pixel_values = pixel_values.repeat_interleave(image_repeat, dim=0)
                pixel_mask = None if pixel_mask is None else pixel_mask.repeat_interleave(image_repeat, dim=0)
                Bi, image_repeat = Bi * image_repeat, 1
            geom, n_slots, sel = self._patch_selection(pixel_mask, Bi, H, W, dev, c.max_image_length, draw_order)
        else:
            geom, n_slots = self._patch_geometry(pixel_mask, Bi, H, W, dev)
        if geom is not None and image_repeat > 1:
            geom = geom.repeat_interleave(image_repeat, dim=0).contiguous()       # the engine reads the geometry per sequence
        pos_rows = self.embeddings.text_embeddings.position_embeddings.weight.shape[0]
        if T > pos_rows:
            raise ValueError(f"text length {T} exceeds the {pos_rows} text position embeddings")
        keep = []

        def as_i64(t):
            if t is None:
                return None
            t = t.to(device=dev, dtype=torch.int64).contiguous()
            keep.append(t)
            return t

        ids, tt, am = as_i64(input_ids), as_i64(token_type_ids), as_i64(attention_mask)
        px = pixel_values.to(dtype=torch.float32).contiguous()
        keep.append(px)
        emb = None
        if inputs_embeds is not None:
            emb = inputs_embeds.detach().to(device=dev, dtype=torch.float32).contiguous()
            keep.append(emb)
        idx_t, idx_s = None, 1
        if isinstance(image_token_type_idx, torch.Tensor):
            idx_t = image_token_type_idx.to(device=dev, dtype=torch.int32).contiguous()
            assert idx_t.numel() == B
            keep.append(idx_t)
        elif image_token_type_idx is not None:
            idx_s = int(image_token_type_idx)
        call = _Call()
        call.keep = keep
        st = self._static_tables()
        call.dims, call.params, call.layers = st["dims"], st["params"], st["layers"]
        n_mod = call.dims.n_modality
        if idx_t is None and not (0 <= idx_s < n_mod):
            raise IndexError(f"image_token_type_idx {idx_s} out of range for {n_mod} modality types")
        E = "embeddings."

        b = _lib.ViltBatchC()
        b.B, b.T, b.H, b.W = B, T, H, W
        b.input_ids, b.inputs_embeds = _lib.ptr(ids), _lib.ptr(emb)
        b.token_type_ids, b.attention_mask = _lib.ptr(tt), _lib.ptr(am)
        b.pixel_values, b.image_type_idx, b.image_type_idx_scalar = _lib.ptr(px), _lib.ptr(idx_t), idx_s
        if geom is not None:
            keep.append(geom)
        b.patch_geom, b.n_patch_slots = _lib.ptr(geom), n_slots
        if sel is not None:
            keep.append(sel)
            b.patch_select = _lib.ptr(sel)
        b.image_repeat = image_repeat
        b.training = int(self.training)
        if self.training and (c.hidden_dropout_prob > 0.0 or c.attention_probs_dropout_prob > 0.0):
            # Philox key of this forward's dropout masks, drawn from torch's CPU generator (seeded by the driver's set_seed,
            # train_upstream_continual_learning.py:103); the backward regenerates the masks from the same key
            b.dropout_seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        call.batch = b
        call.trainable = st["trainable"]
        if emb is not None:      # with inputs_embeds the word table is not on the path (grad stays None)
            call.trainable = [(n, p) for n, p in call.trainable if n != E + "text_embeddings.word_embeddings.weight"]
        call.workspace, call.ws_bytes = None, 0
        return call

    def _patch_geometry(self, pixel_mask, B, H, W, dev):
        """pixel_mask -> (geom int32 [B, 2] on the device, n_patch_slots) for padded batches, or (None, 0) when
        every image fills the grid. As visual_embed (modeling_vilt.py:125-129): the mask is sampled at the patch
        origins (nearest interpolation) and the valid rectangle is read off its first column / row.
        A mask that lives on the HOST (what ViltProcessor returns before `.to(device)`) gives the reference's
        exact sequence length (max_b h_b * w_b slots, :163-170) without any device synchronisation; a mask that
        is already on the GPU is handled without reading it back: all (H/P) * (W/P) slots are kept and the
        padding ones are masked -- same outputs, a few more masked rows."""
        if pixel_mask is None:
            return None, 0
        P = self.config.patch_size
        if tuple(pixel_mask.shape) != (B, H, W):
            raise ValueError(f"pixel_mask shape {tuple(pixel_mask.shape)} does not match pixel_values ({B}, {H}, {W})")
        xm = pixel_mask[:, ::P, ::P]
        h = (xm[:, :, 0] != 0).sum(dim=1)
        w = (xm[:, 0, :] != 0).sum(dim=1)
        geom = torch.stack([h, w], dim=1).to(torch.int32)
        full = (H // P) * (W // P)
        if not pixel_mask.is_cuda:
            n = int((h * w).max())
            if n == full and int((h * w).min()) == full:
                return None, 0                       # no padding anywhere: the fixed-resolution path
            if int((h * w).min()) <= 0:
                raise ValueError("pixel_mask marks an image as entirely padding")
            return geom.to(dev).contiguous(), n
        return geom.contiguous(), full

    def _patch_selection(self, pixel_mask, B, H, W, dev, max_image_length: int, draw_order=None):
        """config.max_image_length > 0, modeling_vilt.py:163-189: the sequence gets min(max_b h_b w_b, max_image_length) patch
        rows; an image with at least that many valid patches keeps a RANDOM subset of them, a smaller image keeps all of its
        patches and is padded with masked rows. The subset is drawn here, on the host, with the reference's own calls in the
        reference's order -- torch.multinomial(torch.ones(v).float(), n) per image on torch's CPU generator (the reference
        builds the weights on the CPU whatever device the model is on), including the draw it spends on choosing padding rows
        -- so that a seeded run keeps exactly the patches the reference keeps. Reading the mask costs one device-to-host
        copy when it lives on the GPU. Returns (geom [B, 2] int32, n_slots, select [B, n_slots] int32: raster index of the
        kept patch in the image's own grid, -1 = padding); select is None when the cap drops nothing."""
        P = self.config.patch_size
        hp, wp = H // P, W // P
        if pixel_mask is None:
            h = torch.full((B,), hp, dtype=torch.int64)
            w = torch.full((B,), wp, dtype=torch.int64)
        else:
            if tuple(pixel_mask.shape) != (B, H, W):
                raise ValueError(f"pixel_mask shape {tuple(pixel_mask.shape)} does not match pixel_values ({B}, {H}, {W})")
            xm = pixel_mask[:, ::P, ::P].cpu()
            h = (xm[:, :, 0] != 0).sum(dim=1)
            w = (xm[:, 0, :] != 0).sum(dim=1)
        eff = h * w
        if int(eff.min()) <= 0:
            raise ValueError("pixel_mask marks an image as entirely padding")
        n = min(int(eff.max()), int(max_image_length))
        geom = torch.stack([h, w], dim=1).to(torch.int32)
        if n == int(eff.max()):
            # nothing is dropped: images at the maximum are permuted (outputs invariant), smaller ones padded = the default path
            if int(eff.min()) == hp * wp:
                return None, 0, None
            return geom.to(dev).contiguous(), n, None
        sel = torch.full((B, n), -1, dtype=torch.int32)
        order = list(range(B)) if draw_order is None else [int(i) for i in draw_order]
        if sorted(order) != list(range(B)):
            raise ValueError("patch_draw_order must be a permutation of the sequences of the call")
        for i in order:
            v = int(eff[i])
            pad = n - v
            if pad <= 0:
                sel[i] = torch.multinomial(torch.ones(v).float(), n).to(torch.int32)          # :177-178
            else:
                torch.multinomial(torch.ones(hp * wp - v).float(), pad, replacement=True)      # :180, keeps the generator in step
                sel[i, :v] = torch.arange(v, dtype=torch.int32)
        return geom.to(dev).contiguous(), n, sel.to(dev).contiguous()

    def _static_tables(self):
        """C structs that only change when the arena is rebuilt, the active adapter changes or a
        requires_grad flag flips: cached between calls."""
        arena = self._arena
        c = self.config
        items = arena.named_items()
        active = self._active_adapter
        from .. import ops
        precision = ops.get_precision()
        key = (id(arena.theta), active, tuple(p.requires_grad for _, p in items), precision, c.hidden_dropout_prob,
               c.attention_probs_dropout_prob)
        cached = getattr(self, "_static_cache", None)
        if cached is not None and cached["key"] == key:
            return cached
        off = arena.offsets
        named = dict(items)
        rg = {n: p.requires_grad for n, p in items}
        spec = self.adapter_specs.get(active) if active else None
        d = _lib.ViltDimsC()
        d.hidden, d.layers, d.heads, d.ffn = c.hidden_size, len(self.encoder.layer), c.num_attention_heads, c.intermediate_size
        d.patch, d.channels = c.patch_size, c.num_channels
        d.pos_grid = int(round(math.sqrt(self.embeddings.position_embeddings.shape[1] - 1)))
        d.n_modality, d.ln_eps = self.embeddings.token_type_embeddings.weight.shape[0], c.layer_norm_eps
        d.vocab_size = self.embeddings.text_embeddings.word_embeddings.weight.shape[0]
        d.type_vocab_size = self.embeddings.text_embeddings.token_type_embeddings.weight.shape[0]
        d.precision = _lib.PREC_BF16X3 if precision == "bf16x3" else _lib.PREC_BF16
        d.hidden_dropout, d.attn_dropout = float(c.hidden_dropout_prob), float(c.attention_probs_dropout_prob)
        layer_flags = [0] * d.layers
        for n in named:
            if n.startswith("encoder.layer.") and rg[n]:
                i = int(n.split(".")[2])
                if ".adapters." in n:
                    if active and f".adapters.{active}." in n:
                        layer_flags[i] |= _lib.TRAIN_ADAPTER
                else:
                    layer_flags[i] |= _lib.TRAIN_BASE
        layers = (_lib.ViltLayerC * d.layers)()
        for i in range(d.layers):
            L = f"encoder.layer.{i}."
            lc = layers[i]
            lc.qkv_w = off[L + "attention.attention.query.weight"]
            lc.qkv_b = off[L + "attention.attention.query.bias"]
            lc.o_w, lc.o_b = off[L + "attention.output.dense.weight"], off[L + "attention.output.dense.bias"]
            lc.fc1_w, lc.fc1_b = off[L + "intermediate.dense.weight"], off[L + "intermediate.dense.bias"]
            lc.fc2_w, lc.fc2_b = off[L + "output.dense.weight"], off[L + "output.dense.bias"]
            lc.ln1_w, lc.ln1_b = off[L + "layernorm_before.weight"], off[L + "layernorm_before.bias"]
            lc.ln2_w, lc.ln2_b = off[L + "layernorm_after.weight"], off[L + "layernorm_after.bias"]
            for site, pre in (("mh", L + f"attention.output.adapters.{active}."), ("out", L + f"output.adapters.{active}.")):
                present = spec is not None and (pre + "adapter_up.weight") in off
                for nm, k2 in (("down_w", "adapter_down.0.weight"), ("down_b", "adapter_down.0.bias"),
                               ("up_w", "adapter_up.weight"), ("up_b", "adapter_up.bias")):
                    setattr(lc, f"{site}_{nm}", off[pre + k2] if present else -1)
            lc.flags = layer_flags[i]
        pc = _lib.ViltParamsC()
        E = "embeddings."
        pc.cls_token, pc.pos_emb = off[E + "cls_token"], off[E + "position_embeddings"]
        pc.word_emb = off[E + "text_embeddings.word_embeddings.weight"]
        pc.text_pos_emb = off[E + "text_embeddings.position_embeddings.weight"]
        pc.text_type_emb = off[E + "text_embeddings.token_type_embeddings.weight"]
        pc.text_ln_w, pc.text_ln_b = off[E + "text_embeddings.LayerNorm.weight"], off[E + "text_embeddings.LayerNorm.bias"]
        pc.patch_w, pc.patch_b = off[E + "patch_embeddings.projection.weight"], off[E + "patch_embeddings.projection.bias"]
        pc.mod_emb = off[E + "token_type_embeddings.weight"]
        pc.final_ln_w, pc.final_ln_b = off["layernorm.weight"], off["layernorm.bias"]
        pc.pooler_w, pc.pooler_b = off["pooler.dense.weight"], off["pooler.dense.bias"]
        pc.layer = ctypes.cast(layers, ctypes.POINTER(_lib.ViltLayerC))
        if spec is not None:
            some = next(n for n in off if f".adapters.{active}.adapter_down.0.bias" in n)
            pc.adapter_r = arena.numels[some]
            pc.adapter_act = _lib.EPI_RELU if spec.non_linearity == "relu" else _lib.EPI_SWISH
        else:
            pc.adapter_r, pc.adapter_act = 0, _lib.EPI_SWISH
        pc.embed_flags = _lib.TRAIN_BASE if any(rg[n] for n in named if n.startswith(E)) else 0
        pc.tail_flags = _lib.TRAIN_BASE if any(rg[n] for n in named if n.startswith(("layernorm.", "pooler."))) else 0
        # which parameters get gradients: everything trainable that the engine touches
        used_adapter = f".adapters.{active}." if active else None
        trainable = [(n, p) for n, p in items
                     if p.requires_grad and (".adapters." not in n or (used_adapter and used_adapter in n))]
        self._static_cache = dict(key=key, dims=d, params=pc, layers=layers, trainable=trainable)
        return self._static_cache

    def _run_forward(self, call: _Call, save: bool) -> torch.Tensor:
        arena = self._arena
        arena.refresh_shadow()
        if call.dims.precision == _lib.PREC_BF16X3:
            call.params.shadow_lo = _lib.ptr(arena.refresh_shadow_lo())
        nbytes = _lib.climb_vilt_forward_workspace_bytes(ctypes.byref(call.dims), ctypes.byref(call.params),
                                                         ctypes.byref(call.batch), int(save))
        if nbytes < 0:
            _lib.check(-1)
        dev = arena.theta.device
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        pooled = torch.empty(call.batch.B, self.config.hidden_size, dtype=torch.float32, device=dev)
        _lib.check(_lib.climb_vilt_forward(ctypes.byref(call.dims), ctypes.byref(call.params), ctypes.byref(call.batch),
                                           _lib.ptr(arena.theta), _lib.ptr(arena.shadow), _lib.ptr(ws), nbytes,
                                           int(save), _lib.ptr(pooled), _lib.stream()))
        if save:
            call.workspace, call.ws_bytes = ws, nbytes
            call.arena_theta = arena.theta
        return pooled

    def _run_backward(self, call: _Call, dpooled: torch.Tensor) -> None:
        arena = self._arena
        if call.arena_theta is not arena.theta:
            raise _lib.ClimbError("parameters were re-allocated between forward and backward")
        if call.dims.precision == _lib.PREC_BF16X3:
            call.params.shadow_lo = _lib.ptr(arena.refresh_shadow_lo())
        nbytes = _lib.climb_vilt_backward_scratch_bytes(ctypes.byref(call.dims), ctypes.byref(call.params),
                                                        ctypes.byref(call.batch))
        if self._scratch is None or self._scratch.numel() < nbytes or self._scratch.device != arena.theta.device:
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=arena.theta.device)
        arena.prepare_grads(call.trainable)
        n_layers = call.dims.layers

        def run(first, last, parts):
            _lib.check(_lib.climb_vilt_backward(ctypes.byref(call.dims), ctypes.byref(call.params), ctypes.byref(call.batch),
                                                _lib.ptr(arena.theta), _lib.ptr(arena.shadow), _lib.ptr(call.workspace),
                                                call.ws_bytes, _lib.ptr(self._scratch), self._scratch.numel(),
                                                _lib.ptr(dpooled), _lib.ptr(arena.grad), first, last, parts, _lib.stream()))

        sync = self.grad_sync
        chunk = getattr(sync, "layers_per_chunk", 0) if sync is not None else 0
        if sync is None or chunk <= 0 or chunk >= n_layers:
            run(n_layers - 1, 0, _lib.BWD_TAIL | _lib.BWD_EMBED)
            arena.publish_grads(call.trainable)
            if sync is not None:
                sync(arena)
        else:
            # data parallel: issue the backward top-down in chunks of layers; after each chunk the
            # gradient span it completed (arena order = network order) is all-reduced on NCCL's stream
            # while the next chunk computes
            off = arena.offsets
            hi_elem = arena.size
            sync.begin(arena)
            chunks = sync.layer_chunks(n_layers) if hasattr(sync, "layer_chunks") else \
                [(f, max(0, f - chunk + 1)) for f in range(n_layers - 1, -1, -chunk)]
            for first, last in chunks:
                run(first, last, _lib.BWD_TAIL if first == n_layers - 1 else 0)
                lo_elem = off[f"encoder.layer.{last}.attention.attention.query.weight"]
                sync.reduce_range(arena, lo_elem, hi_elem)
                hi_elem = lo_elem
            # the embeddings go last and alone: their all-reduce (the 94 MB word table dominates) is the only one
            # with nothing left to hide behind, so the bottom layers' spans are already in flight when it starts
            run(-1, 0, _lib.BWD_EMBED)
            sync.reduce_range(arena, 0, hi_elem)
            arena.publish_grads(call.trainable)
            sync.finish(arena)
        call.workspace = None

    # scratch is a cache, not state
    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_scratch", "grad_sync", "_static_cache") else copy.deepcopy(v, memo)
        return new
