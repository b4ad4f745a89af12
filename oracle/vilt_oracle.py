"""ORACLE (test infrastructure, not product code): a plain-PyTorch fp32 CPU restatement of the
reference arithmetic on CLiMB's ViLT hot path. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; nothing under climb_b200/ does.

It is written functionally over a state dict that uses the reference's own parameter names
(SURVEY.md appendix B), so the same weights drive the reference modules, this oracle and the CUDA
path. Gradients come from torch autograd over these forward functions, exactly as in the reference.

Parity pin: oracle/make_golden.py runs the UNMODIFIED reference (vendored transformers 4.17 fork +
src/modeling/vilt.py + src/cl_algorithms) in this container on the same weights / inputs and stores
its outputs under tests/golden/; tests/test_oracle_golden.py checks this restatement against them
(and against the live reference whenever /root/reference is mounted).

Scope of the restatement: the fixed-resolution path (all images of a batch share one H x W that is
a multiple of the patch size and pixel_mask is all ones). There the reference's random patch
selection (modeling_vilt.py:170-193, torch.multinomial) is a permutation of the patch rows, to
which every output that CLiMB consumes (pooler_output, logits, loss, gradients) is invariant; the
oracle keeps raster order.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class ViltDims:
    """The subset of ViltConfig (configuration_vilt.py:101-124) that shapes the arithmetic."""
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    image_size: int = 384
    patch_size: int = 32
    num_channels: int = 3
    vocab_size: int = 30522
    max_position_embeddings: int = 40
    type_vocab_size: int = 2
    modality_type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    max_image_length: int = -1          # ViltConfig.max_image_length (configuration_vilt.py): > 0 caps the patch rows per image

    @property
    def patch_dim(self) -> int:
        return self.image_size // self.patch_size


# task head specs of src/configs/task_configs.py (num_labels / num_images / model_type only)
TASK_SPECS: Dict[str, Dict] = {
    "vqa": dict(num_labels=3129, num_images=1, model_type="classification"),
    "nlvr2": dict(num_labels=2, num_images=2, model_type="classification"),
    "snli-ve": dict(num_labels=3, num_images=1, model_type="classification"),
    "vcr": dict(num_labels=4, num_images=1, model_type="multi-choice", num_choices=4),
}

ENC = "vilt_encoder.vilt."


# -------------------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by the golden generator, the tests and bench.py)
# -------------------------------------------------------------------------------------------------
def param_shapes(dims: ViltDims, tasks: Sequence[str] = (), adapters: Optional[Dict[str, int]] = None,
                 adapter_sites: Sequence[str] = ("mh", "output"),
                 task_specs: Dict[str, Dict] = TASK_SPECS) -> "OrderedDict[str, Tuple[int, ...]]":
    """Names and shapes of a ViltContinualLearner state dict, in the
    reference's registration order. adapters: {task_key: bottleneck width r}."""
    d, ff, P = dims.hidden_size, dims.intermediate_size, dims.patch_size
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    e = ENC + "embeddings."
    s[e + "cls_token"] = (1, 1, d)
    s[e + "position_embeddings"] = (1, dims.patch_dim ** 2 + 1, d)
    s[e + "text_embeddings.word_embeddings.weight"] = (dims.vocab_size, d)
    s[e + "text_embeddings.position_embeddings.weight"] = (dims.max_position_embeddings, d)
    s[e + "text_embeddings.token_type_embeddings.weight"] = (dims.type_vocab_size, d)
    s[e + "text_embeddings.LayerNorm.weight"] = (d,)
    s[e + "text_embeddings.LayerNorm.bias"] = (d,)
    s[e + "patch_embeddings.projection.weight"] = (d, dims.num_channels, P, P)
    s[e + "patch_embeddings.projection.bias"] = (d,)
    n_types = max(dims.modality_type_vocab_size, 3 if "nlvr2" in tasks else 2)
    s[e + "token_type_embeddings.weight"] = (n_types, d)
    for i in range(dims.num_hidden_layers):
        l = f"{ENC}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[l + f"attention.attention.{n}.weight"] = (d, d)
            s[l + f"attention.attention.{n}.bias"] = (d,)
        s[l + "attention.output.dense.weight"] = (d, d)
        s[l + "attention.output.dense.bias"] = (d,)
        if adapters and "mh" in adapter_sites:
            for t, r in adapters.items():
                a = l + f"attention.output.adapters.{t}."
                s[a + "adapter_down.0.weight"] = (r, d)
                s[a + "adapter_down.0.bias"] = (r,)
                s[a + "adapter_up.weight"] = (d, r)
                s[a + "adapter_up.bias"] = (d,)
        s[l + "intermediate.dense.weight"] = (ff, d)
        s[l + "intermediate.dense.bias"] = (ff,)
        s[l + "output.dense.weight"] = (d, ff)
        s[l + "output.dense.bias"] = (d,)
        if adapters and "output" in adapter_sites:
            for t, r in adapters.items():
                a = l + f"output.adapters.{t}."
                s[a + "adapter_down.0.weight"] = (r, d)
                s[a + "adapter_down.0.bias"] = (r,)
                s[a + "adapter_up.weight"] = (d, r)
                s[a + "adapter_up.bias"] = (d,)
        s[l + "layernorm_before.weight"] = (d,)
        s[l + "layernorm_before.bias"] = (d,)
        s[l + "layernorm_after.weight"] = (d,)
        s[l + "layernorm_after.bias"] = (d,)
    s[ENC + "layernorm.weight"] = (d,)
    s[ENC + "layernorm.bias"] = (d,)
    s[ENC + "pooler.dense.weight"] = (d, d)
    s[ENC + "pooler.dense.bias"] = (d,)
    for t in tasks:
        spec = task_specs[t]
        h = f"task_layer.{t}."
        if spec["model_type"] == "classification":
            s[h + "0.weight"] = (2 * d, d * spec["num_images"])
            s[h + "0.bias"] = (2 * d,)
            s[h + "1.weight"] = (2 * d,)
            s[h + "1.bias"] = (2 * d,)
            s[h + "3.weight"] = (spec["num_labels"], 2 * d)
            s[h + "3.bias"] = (spec["num_labels"],)
        else:
            s[h + "1.weight"] = (1, d)
            s[h + "1.bias"] = (1,)
    return s


def synth_state_dict(dims: ViltDims, tasks: Sequence[str] = (), seed: int = 42,
                     adapters: Optional[Dict[str, int]] = None,
                     adapter_sites: Sequence[str] = ("mh", "output"),
                     layer_scale: float = 1.0, head_scale: float = 1.0) -> "OrderedDict[str, Tensor]":
    """Seeded synthetic weights. Same family as the reference's _init_weights (normal(0, 0.02),
    modeling_vilt.py:597-611) but with non-trivial LayerNorm gains and biases so that every term of
    the arithmetic is exercised. One torch.Generator per tensor: values do not depend on which other
    tensors exist.

    layer_scale / head_scale multiply the weight MATRICES of the encoder layers / the multi-choice head. At the
    0.02 init a tiny encoder passes almost nothing from the other tokens into row 0 ([CLS]), so the four VCR
    choices of a sample (same image, same [CLS] token) pool to nearly the same vector, the loss sits at ln 4 and
    every gradient is a residual of cancelling terms: a fixture that pins nothing. The VCR fixtures therefore
    use layer_scale = 6, head_scale = 20 (information flows, loss != ln 4, gradients well conditioned)."""
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for idx, (name, shape) in enumerate(param_shapes(dims, tasks, adapters, adapter_sites).items()):
        g = torch.Generator().manual_seed(seed * 1_000_003 + _stable_hash(name))
        if name.endswith("LayerNorm.weight") or "layernorm" in name and name.endswith("weight") \
                or (name.startswith("task_layer") and name.endswith("1.weight") and len(shape) == 1):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
            if layer_scale != 1.0 and len(shape) >= 2 and "encoder.layer." in name and ".adapters." not in name:
                t = t * layer_scale
            if head_scale != 1.0 and name.startswith("task_layer.") and len(shape) == 2 and shape[0] == 1:
                t = t * head_scale
        sd[name] = t.float()
    return sd


def _stable_hash(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


# -------------------------------------------------------------------------------------------------
# forward pieces
# -------------------------------------------------------------------------------------------------
def text_embeddings(sd, input_ids: Optional[Tensor], token_type_ids: Optional[Tensor], dims: ViltDims,
                    inputs_embeds: Optional[Tensor] = None) -> Tensor:
    """TextEmbeddings.forward, modeling_vilt.py:272-304 (dropout p = 0 in ViltConfig)."""
    p = ENC + "embeddings.text_embeddings."
    if inputs_embeds is None:
        inputs_embeds = F.embedding(input_ids, sd[p + "word_embeddings.weight"])
    B, T = inputs_embeds.shape[:2]
    if token_type_ids is None:
        token_type_ids = torch.zeros(B, T, dtype=torch.long)
    emb = inputs_embeds + F.embedding(token_type_ids, sd[p + "token_type_embeddings.weight"])
    emb = emb + sd[p + "position_embeddings.weight"][:T][None]
    return F.layer_norm(emb, (dims.hidden_size,), sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"],
                        dims.layer_norm_eps)


def interpolated_position_table(sd, dims: ViltDims, h: int, w: int) -> Tensor:
    """Rows 1.. of position_embeddings bilinearly resized (align_corners=True) from the
    patch_dim x patch_dim training grid to h x w, raster order: modeling_vilt.py:130-147."""
    d = dims.hidden_size
    pos = sd[ENC + "embeddings.position_embeddings"]
    spatial = pos[:, 1:, :].transpose(1, 2).reshape(1, d, dims.patch_dim, dims.patch_dim)
    out = F.interpolate(spatial, size=(h, w), mode="bilinear", align_corners=True)
    return out.flatten(2).transpose(1, 2)[0]          # [h*w, d]


def visual_embed_fixed(sd, pixel_values: Tensor, dims: ViltDims) -> Tensor:
    """ViltEmbeddings.visual_embed (modeling_vilt.py:121-205) on the fixed-resolution path:
    conv patchify (:124, :309-328), interpolated position table (:130-147), [cls] prepended and
    position row 0 added (:195-200). Patch rows stay in raster order (see module docstring)."""
    e = ENC + "embeddings."
    x = F.conv2d(pixel_values, sd[e + "patch_embeddings.projection.weight"],
                 sd[e + "patch_embeddings.projection.bias"], stride=dims.patch_size)
    B, d, h, w = x.shape
    x = x.flatten(2).transpose(1, 2)                                   # [B, h*w, d]
    x = x + interpolated_position_table(sd, dims, h, w)[None]
    cls = sd[e + "cls_token"].expand(B, -1, -1) + sd[e + "position_embeddings"][:, :1, :]
    return torch.cat([cls, x], dim=1)


def patch_geometry(pixel_mask: Tensor, patch_size: int) -> Tuple[Tensor, Tensor]:
    """Valid patch rows / columns per image, modeling_vilt.py:125-129: the pixel mask is resized to the patch grid
    by nearest interpolation (= sampled at the patch origins) and read along its first column / row."""
    xm = pixel_mask[:, ::patch_size, ::patch_size]
    return (xm[:, :, 0] != 0).sum(dim=1), (xm[:, 0, :] != 0).sum(dim=1)


def select_patches(hs: Tensor, ws: Tensor, grid: int, max_image_length: int):
    """modeling_vilt.py:163-189 with max_image_length > 0: n = min(max_b h_b w_b, max_image_length) rows per image; an image
    with v >= n valid patches keeps torch.multinomial(torch.ones(v).float(), n) of them (:177-178, indices into its valid
    patches in raster order), a smaller one keeps all and is padded with n - v masked rows chosen by a second multinomial
    (:180) whose values do not matter but whose draw advances the generator. Same calls, same order as the reference, so
    the same seed gives the same subset. Returns (n, [LongTensor of kept raster indices per image])."""
    eff = hs * ws
    n = min(int(eff.max()), int(max_image_length))
    keep = []
    for b in range(len(hs)):
        v = int(eff[b])
        if n - v <= 0:
            keep.append(torch.multinomial(torch.ones(v).float(), n))
        else:
            torch.multinomial(torch.ones(grid - v).float(), n - v, replacement=True)
            keep.append(torch.arange(v))
    return n, keep


def visual_embed_ragged(sd, pixel_values: Tensor, pixel_mask: Tensor, dims: ViltDims) -> Tuple[Tensor, Tensor]:
    """ViltEmbeddings.visual_embed for a batch padded to a common H x W (modeling_vilt.py:121-205, default
    max_image_length = -1): every image keeps ALL its valid patches with the position table interpolated to
    its own h_b x w_b grid (:132-147); shorter images are padded to max_b h_b * w_b rows whose mask is 0
    (:163-189). The reference permutes the valid rows (multinomial) and fills the padding with randomly chosen
    masked patches; the outputs CLiMB consumes are invariant to both, so the oracle keeps raster order and
    zero padding rows. Returns ([B, 1 + n, d] embeddings, [B, 1 + n] mask)."""
    e = ENC + "embeddings."
    x = F.conv2d(pixel_values, sd[e + "patch_embeddings.projection.weight"],
                 sd[e + "patch_embeddings.projection.bias"], stride=dims.patch_size)
    B, d = x.shape[:2]
    hs, ws = patch_geometry(pixel_mask, dims.patch_size)
    n = int((hs * ws).max())
    keep = None
    if isinstance(dims.max_image_length, int) and 0 < dims.max_image_length < n:
        n, keep = select_patches(hs, ws, x.shape[2] * x.shape[3], dims.max_image_length)
    rows, masks = [], []
    for b in range(B):
        h, w = int(hs[b]), int(ws[b])
        v = x[b, :, :h, :w].flatten(1).transpose(0, 1) + interpolated_position_table(sd, dims, h, w)     # [h*w, d]
        if keep is not None:
            v = v[keep[b]]
        pad = n - v.shape[0]
        rows.append(torch.cat([v, v.new_zeros(pad, d)], dim=0))
        masks.append(torch.cat([torch.ones(v.shape[0]), torch.zeros(pad)]))
    x = torch.stack(rows, dim=0)
    cls = sd[e + "cls_token"].expand(B, -1, -1) + sd[e + "position_embeddings"][:, :1, :]
    return torch.cat([cls, x], dim=1), torch.cat([torch.ones(B, 1), torch.stack(masks, dim=0)], dim=1)


def embeddings(sd, dims: ViltDims, input_ids, attention_mask, token_type_ids, pixel_values,
               image_token_type_idx: int = 1, inputs_embeds=None, pixel_mask=None, masks=None) -> Tuple[Tensor, Tensor]:
    """ViltEmbeddings.forward, modeling_vilt.py:207-246: text || image with modality-type rows.
    masks["embed"] ([B, L, d] keep factors 0 or 1 / (1 - p)) = the two embedding dropouts (:303 on the text LayerNorm output,
    :201 on patches + position embeddings), both applied BEFORE the modality-type rows are added (:231-241)."""
    tt = sd[ENC + "embeddings.token_type_embeddings.weight"]
    text = text_embeddings(sd, input_ids, token_type_ids, dims, inputs_embeds)
    grid = (pixel_values.shape[-2] // dims.patch_size) * (pixel_values.shape[-1] // dims.patch_size)
    capped = isinstance(dims.max_image_length, int) and 0 < dims.max_image_length < grid
    if capped and pixel_mask is None:
        pixel_mask = torch.ones(pixel_values.shape[0], pixel_values.shape[-2], pixel_values.shape[-1], dtype=torch.long)
    if pixel_mask is not None and (capped or not bool((pixel_mask != 0).all())):
        image, image_mask = visual_embed_ragged(sd, pixel_values, pixel_mask, dims)
    else:
        image = visual_embed_fixed(sd, pixel_values, dims)
        image_mask = torch.ones(image.shape[:2], dtype=text.dtype)
    if masks is not None and "embed" in masks:
        T = text.shape[1]
        text = text * masks["embed"][:, :T]
        image = image * masks["embed"][:, T:]
    text = text + tt[0]
    image = image + tt[image_token_type_idx]
    masks = torch.cat([attention_mask.to(text.dtype), image_mask.to(text.dtype)], dim=1)
    return torch.cat([text, image], dim=1), masks


_ACTS = {"swish": F.silu, "relu": F.relu, "gelu": F.gelu}


def adapter_bottleneck(sd, prefix: str, h: Tensor, non_linearity: str, scaling: float = 1.0) -> Tensor:
    """Adapter.forward with the Houlsby / Pfeiffer configs as ViLT calls it (residual input = the
    adapter's own input, "input_tensor" = zeros, no LayerNorm): adapters/modeling.py:120-201,
    mixins/vilt.py:23-69; h + s * (W_u act(W_d h + b_d) + b_u)."""
    down = _ACTS[non_linearity](F.linear(h, sd[prefix + "adapter_down.0.weight"], sd[prefix + "adapter_down.0.bias"]))
    up = F.linear(down, sd[prefix + "adapter_up.weight"], sd[prefix + "adapter_up.bias"])
    return h + scaling * up


@dataclass
class AdapterSpec:
    """Active adapter (Stack[name]) and where it sits. houlsby: mh + output, swish;
    pfeiffer: output only, relu (adapters/configuration.py:259-308)."""
    name: str
    non_linearity: str = "swish"
    sites: Tuple[str, ...] = ("mh", "output")
    scaling: float = 1.0


def vilt_layer(sd, i: int, x: Tensor, ext_mask: Tensor, dims: ViltDims,
               adapter: Optional[AdapterSpec] = None, masks=None) -> Tensor:
    """ViltLayer.forward, modeling_vilt.py:503-525 (pre-LN block). masks (optional, keep factors 0 or 1 / (1 - p)):
    ("attn", i) [B, H, L, L] on the probabilities (:374), ("self_out", i) [B, L, d] on dense(ctx) (:410),
    ("out", i) [B, L, d] on dense(inter) (:482) -- dropout with EXPLICIT masks, so that a test can feed the CUDA path's."""
    masks = masks or {}
    l = f"{ENC}encoder.layer.{i}."
    d, H = dims.hidden_size, dims.num_attention_heads
    dh = d // H
    B, L, _ = x.shape
    h1 = F.layer_norm(x, (d,), sd[l + "layernorm_before.weight"], sd[l + "layernorm_before.bias"], dims.layer_norm_eps)

    def heads(t):                                            # transpose_for_scores, :350-353
        return t.view(B, L, H, dh).permute(0, 2, 1, 3)
    q = heads(F.linear(h1, sd[l + "attention.attention.query.weight"], sd[l + "attention.attention.query.bias"]))
    k = heads(F.linear(h1, sd[l + "attention.attention.key.weight"], sd[l + "attention.attention.key.bias"]))
    v = heads(F.linear(h1, sd[l + "attention.attention.value.weight"], sd[l + "attention.attention.value.bias"]))
    scores = q @ k.transpose(-1, -2) / math.sqrt(dh) + ext_mask       # :363-368
    probs = torch.softmax(scores, dim=-1)                              # :371
    if ("attn", i) in masks:
        probs = probs * masks[("attn", i)]                             # :374-375
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, L, d)             # :381-384
    a = F.linear(ctx, sd[l + "attention.output.dense.weight"], sd[l + "attention.output.dense.bias"])   # :409
    if ("self_out", i) in masks:
        a = a * masks[("self_out", i)]                                 # :410
    if adapter is not None and "mh" in adapter.sites:
        a = adapter_bottleneck(sd, l + f"attention.output.adapters.{adapter.name}.", a, adapter.non_linearity, adapter.scaling)
    x = a + x                                                          # :514
    h2 = F.layer_norm(x, (d,), sd[l + "layernorm_after.weight"], sd[l + "layernorm_after.bias"], dims.layer_norm_eps)
    inter = F.gelu(F.linear(h2, sd[l + "intermediate.dense.weight"], sd[l + "intermediate.dense.bias"]))  # :461-466
    o = F.linear(inter, sd[l + "output.dense.weight"], sd[l + "output.dense.bias"])                       # :481
    if ("out", i) in masks:
        o = o * masks[("out", i)]                                      # :482
    o = o + x                                                          # :484
    if adapter is not None and "output" in adapter.sites:
        o = adapter_bottleneck(sd, l + f"output.adapters.{adapter.name}.", o, adapter.non_linearity, adapter.scaling)
    return o


def vilt_forward(sd, dims: ViltDims, input_ids, attention_mask, token_type_ids, pixel_values,
                 image_token_type_idx: int = 1, inputs_embeds=None,
                 adapter: Optional[AdapterSpec] = None, return_hidden: bool = False, pixel_mask=None, dropout_masks=None):
    """ViltModel.forward -> pooler_output, modeling_vilt.py:777-884, 887-899."""
    x, masks = embeddings(sd, dims, input_ids, attention_mask, token_type_ids, pixel_values,
                          image_token_type_idx, inputs_embeds, pixel_mask, masks=dropout_masks)
    ext = (1.0 - masks)[:, None, None, :] * -10000.0                   # modeling_utils.py:299-311
    for i in range(dims.num_hidden_layers):
        x = vilt_layer(sd, i, x, ext, dims, adapter, masks=dropout_masks)
    x = F.layer_norm(x, (dims.hidden_size,), sd[ENC + "layernorm.weight"], sd[ENC + "layernorm.bias"], dims.layer_norm_eps)
    pooled = torch.tanh(F.linear(x[:, 0], sd[ENC + "pooler.dense.weight"], sd[ENC + "pooler.dense.bias"]))
    return (pooled, x) if return_hidden else pooled


def task_head(sd, task: str, pooled: Tensor, spec: Optional[Dict] = None, train: bool = False) -> Tensor:
    """Task heads of ViltContinualLearner.add_task_layer, src/modeling/vilt.py:179-203.
    classification: Linear -> LayerNorm(eps 1e-5) -> GELU -> Linear. multi-choice: Dropout(0.1) ->
    Linear(d, 1) on [B, choices, d], squeezed (:349). Dropout is identity here (eval / p handled by
    the caller): the oracle is deterministic."""
    spec = spec or TASK_SPECS[task]
    h = f"task_layer.{task}."
    if spec["model_type"] == "classification":
        z = F.linear(pooled, sd[h + "0.weight"], sd[h + "0.bias"])
        z = F.gelu(F.layer_norm(z, (z.shape[-1],), sd[h + "1.weight"], sd[h + "1.bias"], 1e-5))
        return F.linear(z, sd[h + "3.weight"], sd[h + "3.bias"])
    return F.linear(pooled, sd[h + "1.weight"], sd[h + "1.bias"]).squeeze()


def learner_forward(sd, dims: ViltDims, task: str, batch: Dict[str, Tensor],
                    adapter: Optional[AdapterSpec] = None, spec: Optional[Dict] = None, dropout_masks=None):
    """ViltContinualLearner.forward on tensor inputs (src/modeling/vilt.py:218-350):
    single image (:241-261), NLVR2 two passes with image_token_type_idx 1, 2 (:291-304), VCR four
    text choices over the same pixels (:334-349). Returns (pooled, logits)."""
    spec = spec or TASK_SPECS[task]
    ids, am, tt, px = batch["input_ids"], batch["attention_mask"], batch["token_type_ids"], batch["pixel_values"]
    pm = batch.get("pixel_mask")          # [B, (n_img,) H, W] or absent (= all ones)
    if spec["model_type"] == "multi-choice":
        outs = [vilt_forward(sd, dims, ids[:, c], am[:, c], tt[:, c], px, 1, adapter=adapter, pixel_mask=pm)
                for c in range(spec["num_choices"])]
        pooled = torch.stack(outs, dim=0).transpose(0, 1)
    elif spec["num_images"] == 1:
        pooled = vilt_forward(sd, dims, ids, am, tt, px, 1, adapter=adapter, pixel_mask=pm, dropout_masks=dropout_masks)
    else:
        outs = [vilt_forward(sd, dims, ids, am, tt, px[:, i], i + 1, adapter=adapter,
                             pixel_mask=None if pm is None else pm[:, i])
                for i in range(spec["num_images"])]
        pooled = torch.cat(outs, dim=-1)
    return pooled, task_head(sd, task, pooled, spec)


def pad_batch_images(batch: Dict[str, Tensor], sizes: Sequence[Tuple[int, int]]) -> Dict[str, Tensor]:
    """Turn a fixed-resolution synthetic batch into what ViltFeatureExtractor returns for images of different
    sizes (feature_extraction_vilt.py:253-292): image k keeps its top-left sizes[k] = (H_k, W_k) pixels, the
    rest is zero padding with pixel_mask = 0. sizes runs over the flattened images (B * n_img)."""
    px = batch["pixel_values"]
    flat = px.reshape(-1, *px.shape[-3:]).clone()
    mask = torch.zeros(flat.shape[0], flat.shape[-2], flat.shape[-1], dtype=torch.long)
    assert len(sizes) == flat.shape[0]
    for k, (hk, wk) in enumerate(sizes):
        flat[k, :, hk:, :] = 0
        flat[k, :, :, wk:] = 0
        mask[k, :hk, :wk] = 1
    out = dict(batch)
    out["pixel_values"] = flat.reshape(px.shape)
    out["pixel_mask"] = mask.reshape(*px.shape[:-3], *px.shape[-2:])
    return out


def task_loss(task: str, logits: Tensor, target: Tensor) -> Tensor:
    """VQA: BCEWithLogits(mean) * num_labels on soft scores (train_vqa.py:95,157);
    others: CrossEntropyLoss (train_nlvr2.py:80,133; train_snli_ve.py; train_vcr.py:83,135)."""
    if task == "vqa":
        return F.binary_cross_entropy_with_logits(logits, target) * target.shape[1]
    return F.cross_entropy(logits, target)


# -------------------------------------------------------------------------------------------------
# EWC (src/cl_algorithms/ewc.py)
# -------------------------------------------------------------------------------------------------
def ewc_penalty(params: Dict[str, Tensor], theta_star: Dict[str, Tensor], fisher: Dict[str, Tensor],
                ewc_loss_weight: float) -> Tensor:
    """EWC.compute_ewc_loss, ewc.py:75-87: lambda * sum_p sum(F_p * (theta_p - theta*_p)^2) over the
    encoder's named parameters."""
    loss = torch.zeros(())
    for name, p in params.items():
        loss = loss + (fisher[name] * (p - theta_star[name]) ** 2).sum()
    return ewc_loss_weight * loss


def fisher_from_batch_grads(batch_grads: List[Dict[str, Tensor]], batch_sizes: List[int]) -> Dict[str, Tensor]:
    """EWC.save_task_parameters, ewc.py:55-71, including its quirk: gradients are never zeroed inside
    the loop, so batch t contributes (sum_{s<=t} g_s)^2, and the sum is divided by the sample count."""
    fisher: Dict[str, Tensor] = {}
    running: Dict[str, Tensor] = {}
    for g in batch_grads:
        for n, t in g.items():
            running[n] = running.get(n, torch.zeros_like(t)) + t
            fisher[n] = fisher.get(n, torch.zeros_like(t)) + running[n] ** 2
    total = float(sum(batch_sizes))
    return {n: t / total for n, t in fisher.items()}


# -------------------------------------------------------------------------------------------------
# optimizer grouping (src/modeling/vilt.py:205-215) and LR schedule (optimization.py:208-220)
# -------------------------------------------------------------------------------------------------
def weight_decay_groups(names: Sequence[str]) -> Tuple[List[str], List[str]]:
    no_decay = ["bias", "LayerNorm.weight"]
    decay = [n for n in names if not any(nd in n for nd in no_decay)]
    nodecay = [n for n in names if any(nd in n for nd in no_decay)]
    return decay, nodecay


def linear_warmup_decay(step: int, warmup: int, total: int) -> float:
    """get_polynomial_decay_schedule_with_warmup(power=1, lr_end=0) multiplier (train_vqa.py:199-205)."""
    if step < warmup:
        return step / max(1, warmup)
    if step > total:
        return 0.0
    return 1.0 - (step - warmup) / (total - warmup)


# -------------------------------------------------------------------------------------------------
# synthetic batches
# -------------------------------------------------------------------------------------------------
def synth_batch(task: str, B: int, dims: ViltDims, T: int = 40, image_hw: Tuple[int, int] = (448, 448),
                seed: int = 0, masked: bool = False) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(10_000 + seed)
    spec = TASK_SPECS[task]
    n_choices = spec.get("num_choices", 1)
    lo, hi = min(1000, dims.vocab_size // 2), dims.vocab_size
    ids = torch.randint(lo, hi, (B, n_choices, T), generator=g)
    ids[..., 0] = min(101, dims.vocab_size - 2)
    am = torch.ones(B, n_choices, T, dtype=torch.long)
    if masked:
        lens = torch.randint(max(2, T // 5), T + 1, (B, n_choices), generator=g)
        am = (torch.arange(T)[None, None, :] < lens[..., None]).long()
        ids = ids * am
    last = am.sum(-1) - 1
    ids.scatter_(-1, last[..., None], min(102, dims.vocab_size - 1))
    tt = torch.zeros(B, n_choices, T, dtype=torch.long)
    H, W = image_hw
    n_img = spec["num_images"]
    px = torch.rand(B, n_img, dims.num_channels, H, W, generator=g) * 2 - 1
    batch = dict(input_ids=ids, attention_mask=am, token_type_ids=tt, pixel_values=px)
    if n_choices == 1:
        for k in ("input_ids", "attention_mask", "token_type_ids"):
            batch[k] = batch[k][:, 0]
    if n_img == 1:
        batch["pixel_values"] = px[:, 0]
    if task == "vqa":
        tgt = torch.zeros(B, spec["num_labels"])
        for b in range(B):
            k = int(torch.randint(1, 4, (1,), generator=g))
            idx = torch.randperm(spec["num_labels"], generator=g)[:k]
            tgt[b, idx] = torch.tensor([0.3, 0.6, 0.9, 1.0])[torch.randint(0, 4, (k,), generator=g)]
        batch["target"] = tgt
    else:
        batch["target"] = torch.randint(0, spec["num_labels"], (B,), generator=g)
    return batch


# -------------------------------------------------------------------------------------------------
# ViLT-BERT (src/modeling/viltbert.py): frozen BertModel -> inputs_embeds of ViltModel
# -------------------------------------------------------------------------------------------------
@dataclass
class BertDims:
    """The subset of BertConfig that shapes the arithmetic (defaults = bert-base-uncased)."""
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    vocab_size: int = 30522
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12


VB_ENC = "viltbert_encoder.vilt."
VB_BERT = "viltbert_encoder.bert."


def bert_param_shapes(b: BertDims, prefix: str = VB_BERT) -> "OrderedDict[str, Tuple[int, ...]]":
    """Names / shapes of adapter-transformers' BertModel (modeling_bert.py:171-192, 231-262, 362-370, 430-453,
    646-650), registration order."""
    d, ff = b.hidden_size, b.intermediate_size
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    e = prefix + "embeddings."
    s[e + "word_embeddings.weight"] = (b.vocab_size, d)
    s[e + "position_embeddings.weight"] = (b.max_position_embeddings, d)
    s[e + "token_type_embeddings.weight"] = (b.type_vocab_size, d)
    s[e + "LayerNorm.weight"] = (d,)
    s[e + "LayerNorm.bias"] = (d,)
    for i in range(b.num_hidden_layers):
        l = f"{prefix}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[l + f"attention.self.{n}.weight"] = (d, d)
            s[l + f"attention.self.{n}.bias"] = (d,)
        s[l + "attention.output.dense.weight"] = (d, d)
        s[l + "attention.output.dense.bias"] = (d,)
        s[l + "attention.output.LayerNorm.weight"] = (d,)
        s[l + "attention.output.LayerNorm.bias"] = (d,)
        s[l + "intermediate.dense.weight"] = (ff, d)
        s[l + "intermediate.dense.bias"] = (ff,)
        s[l + "output.dense.weight"] = (d, ff)
        s[l + "output.dense.bias"] = (d,)
        s[l + "output.LayerNorm.weight"] = (d,)
        s[l + "output.LayerNorm.bias"] = (d,)
    s[prefix + "pooler.dense.weight"] = (d, d)
    s[prefix + "pooler.dense.bias"] = (d,)
    return s


def synth_bert_state_dict(b: BertDims, seed: int = 42, prefix: str = VB_BERT) -> "OrderedDict[str, Tensor]":
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for name, shape in bert_param_shapes(b, prefix).items():
        g = torch.Generator().manual_seed(seed * 1_000_003 + _stable_hash(name))
        if name.endswith("LayerNorm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t.float()
    return sd


def synth_viltbert_state_dict(dims: ViltDims, b: BertDims, tasks: Sequence[str] = (), seed: int = 42,
                              layer_scale: float = 1.0, head_scale: float = 1.0):
    """ViltBertContinualLearner state dict: the ViLT learner's tensors under `viltbert_encoder.vilt.` (same
    values as synth_state_dict: the generator is keyed by the ViLT-learner name) + BERT + task heads."""
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for k, v in synth_state_dict(dims, tasks, seed, layer_scale=layer_scale, head_scale=head_scale).items():
        sd[VB_ENC + k[len(ENC):] if k.startswith(ENC) else k] = v
    sd.update(synth_bert_state_dict(b, seed))
    return sd


def bert_forward(sd, b: BertDims, input_ids: Tensor, attention_mask: Optional[Tensor], token_type_ids: Optional[Tensor],
                 prefix: str = VB_BERT) -> Tensor:
    """BertModel.forward -> last_hidden_state in eval mode (dropout off): modeling_bert.py:918-1054;
    BertEmbeddings :194-228, BertSelfAttention :265-360, BertSelfOutput :372-376 (post-LN),
    BertIntermediate :439-442, BertOutput :455-459."""
    d, H = b.hidden_size, b.num_attention_heads
    dh = d // H
    B, T = input_ids.shape
    e = prefix + "embeddings."
    if token_type_ids is None:
        token_type_ids = torch.zeros(B, T, dtype=torch.long)
    x = F.embedding(input_ids, sd[e + "word_embeddings.weight"]) + F.embedding(token_type_ids, sd[e + "token_type_embeddings.weight"])
    x = x + sd[e + "position_embeddings.weight"][:T][None]
    x = F.layer_norm(x, (d,), sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], b.layer_norm_eps)
    if attention_mask is None:
        attention_mask = torch.ones(B, T, dtype=torch.long)
    ext = (1.0 - attention_mask.to(x.dtype))[:, None, None, :] * -10000.0          # modeling_utils.py:299-311
    for i in range(b.num_hidden_layers):
        l = f"{prefix}encoder.layer.{i}."
        heads = lambda t: t.view(B, T, H, dh).permute(0, 2, 1, 3)
        q = heads(F.linear(x, sd[l + "attention.self.query.weight"], sd[l + "attention.self.query.bias"]))
        k = heads(F.linear(x, sd[l + "attention.self.key.weight"], sd[l + "attention.self.key.bias"]))
        v = heads(F.linear(x, sd[l + "attention.self.value.weight"], sd[l + "attention.self.value.bias"]))
        probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + ext, dim=-1)
        ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, T, d)
        a = F.linear(ctx, sd[l + "attention.output.dense.weight"], sd[l + "attention.output.dense.bias"])
        x = F.layer_norm(a + x, (d,), sd[l + "attention.output.LayerNorm.weight"], sd[l + "attention.output.LayerNorm.bias"],
                         b.layer_norm_eps)
        inter = F.gelu(F.linear(x, sd[l + "intermediate.dense.weight"], sd[l + "intermediate.dense.bias"]))
        o = F.linear(inter, sd[l + "output.dense.weight"], sd[l + "output.dense.bias"])
        x = F.layer_norm(o + x, (d,), sd[l + "output.LayerNorm.weight"], sd[l + "output.LayerNorm.bias"], b.layer_norm_eps)
    return x


def viltbert_learner_forward(sd, dims: ViltDims, b: BertDims, task: str, batch: Dict[str, Tensor], spec: Optional[Dict] = None):
    """ViltBertContinualLearner.forward on tensor inputs (src/modeling/viltbert.py:231-345): every encoder pass
    first runs the frozen BERT under no_grad (:115-120, :145) and feeds its last_hidden_state to ViltModel as
    inputs_embeds with input_ids = None (:146-149). Returns (pooled, logits)."""
    spec = spec or TASK_SPECS[task]
    vsd = {(ENC + k[len(VB_ENC):] if k.startswith(VB_ENC) else k): v for k, v in sd.items()}      # same tensors, ViLT names
    ids, am, tt, px = batch["input_ids"], batch["attention_mask"], batch["token_type_ids"], batch["pixel_values"]

    def enc(ids_, am_, tt_, px_, idx):
        with torch.no_grad():
            feats = bert_forward(sd, b, ids_, am_, tt_)
        return vilt_forward(vsd, dims, None, am_, tt_, px_, idx, inputs_embeds=feats)

    if spec["model_type"] == "multi-choice":
        outs = [enc(ids[:, c], am[:, c], tt[:, c], px, 1) for c in range(spec["num_choices"])]
        pooled = torch.stack(outs, dim=0).transpose(0, 1)
    elif spec["num_images"] == 1:
        pooled = enc(ids, am, tt, px, 1)
    else:
        outs = [enc(ids, am, tt, px[:, i], i + 1) for i in range(spec["num_images"])]
        pooled = torch.cat(outs, dim=-1)
    return pooled, task_head(vsd, task, pooled, spec)
