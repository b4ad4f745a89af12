"""B200BertModel: the frozen text encoder of CLiMB's ViLT-BERT (src/modeling/viltbert.py:115-120,
`BertModel.from_pretrained("bert-base-uncased")` run under torch.no_grad()).

Same parameter tree (names, shapes, registration order) as adapter-transformers'
`transformers.models.bert.modeling_bert.BertModel` (modeling_bert.py:869-1054) so that
`bert-base-uncased` weights and CLiMB ViLT-BERT checkpoints (`viltbert_encoder.bert.*`) load
unchanged; forward is ONE call into libclimb_b200.so (include/climb_b200.h: climb_bert_forward) and is
forward-only: the reference never differentiates through BERT, so `last_hidden_state` comes back
detached. The nn.Linear / nn.LayerNorm / nn.Embedding children are parameter containers only.

Reference quirk kept: BertConfig's hidden_dropout_prob = attention_probs_dropout_prob = 0.1 stay active
inside no_grad whenever the module is in train mode (viltbert.py never puts `bert` in eval mode during
training), so train-mode features are stochastic. The masks here come from a counter-based generator
seeded from torch's default generator (`torch.manual_seed` reproduces a run; the bit pattern differs from
ATen's Philox offsets, the distribution does not).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ..arena import ParamArena


@dataclass
class B200BertConfig:
    """Fields of BertConfig (configuration_bert.py) that the forward reads; defaults = bert-base-uncased."""
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    initializer_range: float = 0.02
    layer_norm_eps: float = 1e-12
    pad_token_id: int = 0
    position_embedding_type: str = "absolute"

    @classmethod
    def from_hf(cls, cfg) -> "B200BertConfig":
        get = (lambda k, dflt: cfg.get(k, dflt)) if isinstance(cfg, dict) else (lambda k, dflt: getattr(cfg, k, dflt))
        return cls(**{f: get(f, getattr(cls, f)) for f in cls.__dataclass_fields__})


class _BertEmbeddings(nn.Module):
    def __init__(self, c: B200BertConfig):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size, padding_idx=c.pad_token_id)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(c.max_position_embeddings).expand((1, -1)))


class _BertSelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.query = nn.Linear(c.hidden_size, c.hidden_size)
        self.key = nn.Linear(c.hidden_size, c.hidden_size)
        self.value = nn.Linear(c.hidden_size, c.hidden_size)


class _BertSelfOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _BertAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = _BertSelfAttention(c)
        self.output = _BertSelfOutput(c)


class _BertIntermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)


class _BertOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.intermediate_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _BertAttention(c)
        self.intermediate = _BertIntermediate(c)
        self.output = _BertOutput(c)


class _BertEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(c) for _ in range(c.num_hidden_layers)])


class _BertPooler(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)


@dataclass
class BertOutput:
    """BaseModelOutputWithPoolingAndCrossAttentions reduced to what ViLT-BERT reads (viltbert.py:120)."""
    last_hidden_state: torch.Tensor
    pooler_output: Optional[torch.Tensor] = None

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class B200BertModel(nn.Module):
    def __init__(self, config=None, add_pooling_layer: bool = True):
        super().__init__()
        self.config = config if isinstance(config, B200BertConfig) else B200BertConfig.from_hf(config or {})
        c = self.config
        if c.hidden_size != c.num_attention_heads * 64 or c.hidden_size % 128:
            raise ValueError("climb_b200 kernels need head_dim 64 and hidden_size % 128 == 0 "
                             f"(got hidden={c.hidden_size}, heads={c.num_attention_heads})")
        if c.hidden_act != "gelu" or c.position_embedding_type != "absolute":
            raise NotImplementedError("only hidden_act='gelu' (erf) and absolute position embeddings are implemented")
        self.embeddings = _BertEmbeddings(c)
        self.encoder = _BertEncoder(c)
        self.pooler = _BertPooler(c) if add_pooling_layer else None     # kept for checkpoint keys; never evaluated
        self._arena = ParamArena(self, "_arena_items", with_grad=False)
        self.apply(self._init_weights)

    def _init_weights(self, m):          # BertPreTrainedModel._init_weights, modeling_bert.py:744-757
        std = self.config.initializer_range
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)

    def _arena_items(self) -> List[Tuple[str, nn.Parameter]]:
        """Network order; q, k, v weights (and biases) of a layer adjacent so that one [3d, d] GEMM reads them."""
        named = dict(self.named_parameters())
        order = [n for n in named if n.startswith("embeddings.")]
        for i in range(len(self.encoder.layer)):
            pre = f"encoder.layer.{i}."
            a = pre + "attention.self."
            qkv = [a + "query.weight", a + "key.weight", a + "value.weight", a + "query.bias", a + "key.bias", a + "value.bias"]
            order += qkv
            skip = set(qkv)
            order += [n for n in named if n.startswith(pre) and n not in skip]
        seen = set(order)
        order += [n for n in named if n not in seen]
        return [(n, named[n]) for n in order]

    def _tables(self):
        arena = self._arena
        key = id(arena.theta)
        cached = getattr(self, "_static_cache", None)
        if cached is not None and cached["key"] == key:
            return cached
        c, off = self.config, arena.offsets
        d = _lib.BertDimsC()
        d.hidden, d.layers, d.heads, d.ffn, d.ln_eps = (c.hidden_size, len(self.encoder.layer), c.num_attention_heads,
                                                        c.intermediate_size, c.layer_norm_eps)
        layers = (_lib.BertLayerC * d.layers)()
        for i in range(d.layers):
            L = f"encoder.layer.{i}."
            lc = layers[i]
            lc.qkv_w, lc.qkv_b = off[L + "attention.self.query.weight"], off[L + "attention.self.query.bias"]
            lc.o_w, lc.o_b = off[L + "attention.output.dense.weight"], off[L + "attention.output.dense.bias"]
            lc.attn_ln_w, lc.attn_ln_b = off[L + "attention.output.LayerNorm.weight"], off[L + "attention.output.LayerNorm.bias"]
            lc.fc1_w, lc.fc1_b = off[L + "intermediate.dense.weight"], off[L + "intermediate.dense.bias"]
            lc.fc2_w, lc.fc2_b = off[L + "output.dense.weight"], off[L + "output.dense.bias"]
            lc.out_ln_w, lc.out_ln_b = off[L + "output.LayerNorm.weight"], off[L + "output.LayerNorm.bias"]
        pc = _lib.BertParamsC()
        E = "embeddings."
        pc.word_emb, pc.pos_emb = off[E + "word_embeddings.weight"], off[E + "position_embeddings.weight"]
        pc.type_emb = off[E + "token_type_embeddings.weight"]
        pc.emb_ln_w, pc.emb_ln_b = off[E + "LayerNorm.weight"], off[E + "LayerNorm.bias"]
        pc.layer = ctypes.cast(layers, ctypes.POINTER(_lib.BertLayerC))
        self._static_cache = dict(key=key, dims=d, params=pc, layers=layers)
        return self._static_cache

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, **unused):
        """BertModel.forward (modeling_bert.py:918-1054) -> last_hidden_state [B, T, hidden], detached."""
        if input_ids is None or inputs_embeds is not None or position_ids is not None or head_mask is not None:
            raise NotImplementedError("climb_b200's BERT runs ViLT-BERT's call: input_ids (+ attention_mask, token_type_ids)")
        if not input_ids.is_cuda:
            raise _lib.ClimbError("climb_b200 runs on CUDA tensors only (no CPU fallback)")
        dev = input_ids.device
        arena = self._arena
        arena.sync(dev)
        arena.refresh_shadow()
        st = self._tables()
        B, T = input_ids.shape
        if T > self.embeddings.position_embeddings.weight.shape[0]:
            raise ValueError(f"text length {T} exceeds BERT's {self.embeddings.position_embeddings.weight.shape[0]} positions")
        as_i64 = lambda t: None if t is None else t.to(device=dev, dtype=torch.int64).contiguous()
        ids, am, tt = as_i64(input_ids), as_i64(attention_mask), as_i64(token_type_ids)
        b = _lib.BertBatchC()
        b.B, b.T = B, T
        b.input_ids, b.token_type_ids, b.attention_mask = _lib.ptr(ids), _lib.ptr(tt), _lib.ptr(am)
        nbytes = _lib.climb_bert_forward_workspace_bytes(ctypes.byref(st["dims"]), ctypes.byref(b))
        if nbytes < 0:
            _lib.check(-1)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty(B, T, self.config.hidden_size, dtype=torch.float32, device=dev)
        p_hid = float(self.config.hidden_dropout_prob) if self.training else 0.0
        p_att = float(self.config.attention_probs_dropout_prob) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,), device="cpu").item()) if (p_hid > 0 or p_att > 0) else 0
        _lib.check(_lib.climb_bert_forward(ctypes.byref(st["dims"]), ctypes.byref(st["params"]), ctypes.byref(b),
                                           _lib.ptr(arena.theta), _lib.ptr(arena.shadow), _lib.ptr(ws), nbytes,
                                           p_hid, p_att, seed, _lib.ptr(out), _lib.stream()))
        return BertOutput(last_hidden_state=out)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_static_cache" else copy.deepcopy(v, memo)
        return new
