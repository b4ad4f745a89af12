"""B200 drop-ins for CLiMB's ViLT wrappers (src/modeling/vilt.py): same class surface, same
registry signatures, same state-dict keys -- the encoder arithmetic runs in libclimb_b200.so.

    ViltEncoderWrapper      -> B200ViltEncoderWrapper      (vilt.py:30-144)
    ViltContinualLearner    -> B200ViltContinualLearner    (vilt.py:147-367)
    load_vilt_encoder       -> load_vilt_encoder           (vilt.py:481-514)
    create_vilt_continual_learner_model                    (vilt.py:516-546)
    convert_batch_to_vilt_input_dict / convert_seq_batch_to_vilt_input_dict (vilt.py:548-567)

Differences that are deliberate:
  * NLVR2's two passes and VCR's four passes (vilt.py:291-304, 334-347) are batched into ONE encoder
    call over 2B / 4B sequences -- numerically the same rows, no cross-row op exists in the encoder;
  * `forward_tensors(task_key, encodings)` takes the processor's output directly (the synthetic
    benchmark and the tests feed tensors; the PIL / tokenizer path of process_inputs still works
    when a processor is supplied).
"""
from __future__ import annotations

import itertools
import logging
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..optim import ArenaAdamW
from .continual_learner import ContinualLearner, EncoderWrapper
from .vilt_model import B200ViltConfig, B200ViltModel

logger = logging.getLogger(__name__)


class B200ViltEncoderWrapper(EncoderWrapper):
    def __init__(self, processor, vilt: B200ViltModel, device: torch.device):
        super().__init__()
        self.processor = processor
        self.vilt = vilt
        self.device = device
        self.max_text_length = self.vilt.config.max_position_embeddings
        self.encoder_dim = self.vilt.config.hidden_size
        self.gpu_image_pipeline = True      # process_inputs: image half of the ViltProcessor on the GPU (False = the reference's PIL path)
        self.native_tokenizer = True        # process_inputs: BERT WordPiece in native host code (False = the processor's own tokenizer)
        self._image_fe, self._image_fe_key = None, None
        self._native_tok, self._native_tok_key = None, None

    def reset_processor(self, max_text_length: int, img_size: tuple):
        self.max_text_length = max_text_length
        if self.processor is not None:
            self.processor.feature_extractor.size = img_size

    def reallocate_text_image(self, pretrained_pos_emb: torch.Tensor, max_len: int, img_size: int):
        """vilt.py:57-81: tile the text position table so longer language inputs fit."""
        cfg = self.vilt.config
        assert max_len % cfg.max_position_embeddings == 0
        self.reset_processor(max_len, img_size)
        extended = torch.cat([pretrained_pos_emb for _ in range(0, max_len, cfg.max_position_embeddings)], 0)
        te = self.vilt.embeddings.text_embeddings
        te.position_embeddings = nn.Embedding(max_len, cfg.hidden_size).from_pretrained(extended, freeze=False)
        te.register_buffer("position_ids", torch.arange(max_len).expand((1, -1)))

    def tokenize(self, tok, texts: List[str]) -> Dict:
        """The tokenizer call of ViltProcessor.__call__ (processing_vilt.py:72-89) as process_inputs makes it
        (vilt.py:93-95: padding=True, truncation=True, max_length). A transformers BertTokenizer(Fast) is replaced by the
        native WordPiece pipeline over its own vocabulary / casing (climb_b200/text_processing.py: identical ids, rows
        written into pinned staging memory); anything else is called as is."""
        # (tokens ADDED on top of the vocabulary are matched in the raw text by the library; only BERT's own specials are here)
        # The native path covers what process_inputs sends on the hot path: one string per sample through a BERT tokenizer
        # with default normalisation. Text PAIRS (convert_mc_batch_to_vilt_input_dict, vilt.py:561-567: [[a, b], ...] ->
        # token_type_ids 1 on the second segment, pair truncation) and tokenizers built with non-default strip_accents /
        # never_split keep the processor's own tokenizer.
        init = getattr(tok, "init_kwargs", {}) or {}
        plain = all(isinstance(t, str) for t in texts)
        default_norm = init.get("strip_accents", None) is None and not init.get("never_split", None) \
            and getattr(tok, "strip_accents", None) is None
        if (self.native_tokenizer and plain and default_norm
                and type(tok).__name__ in ("BertTokenizerFast", "BertTokenizer") and hasattr(tok, "get_vocab")
                and set(getattr(tok, "get_added_vocab", dict)()) <= {"[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]"}):
            if self._native_tok is None or self._native_tok_key != id(tok):
                from ..text_processing import B200BertTokenizer
                self._native_tok, self._native_tok_key = B200BertTokenizer.from_hf(tok), id(tok)
            return self._native_tok(texts, max_length=self.max_text_length, padding=True, truncation=True,
                                    pin_memory=torch.cuda.is_available())
        return tok(text=texts, max_length=self.max_text_length, padding=True, truncation=True, return_tensors='pt')

    def process_inputs(self, images: List, texts: List[str]) -> Dict:
        if self.processor is None:
            raise RuntimeError("this encoder was built without a ViltProcessor: call forward_tensors() / pass "
                               "encodings, or construct it with processor=ViltProcessor.from_pretrained(...)")
        fe = getattr(self.processor, "feature_extractor", None)
        tok = getattr(self.processor, "tokenizer", None)
        if (self.gpu_image_pipeline and fe is not None and tok is not None and torch.device(self.device).type == "cuda"
                and getattr(fe, "do_resize", True) and getattr(fe, "do_normalize", True) and int(getattr(fe, "resample", 3)) == 3):
            # ViltProcessor.__call__ = tokenizer(text) + feature_extractor(images) (processing_vilt.py:63-107): the text half stays
            # on the host tokenizer, the image half (PIL bicubic resize, normalise, pad, pixel mask: the reference's CPU
            # bottleneck) runs on the GPU with the extractor's own parameters, bit-exact (climb_b200/image_processing.py)
            from ..image_processing import B200ViltFeatureExtractor
            key = (fe.size, getattr(fe, "size_divisor", 32), tuple(fe.image_mean), tuple(fe.image_std))
            if self._image_fe is None or self._image_fe_key != key:
                self._image_fe = B200ViltFeatureExtractor(size=key[0], size_divisor=key[1], image_mean=key[2], image_std=key[3],
                                                          device=self.device)
                self._image_fe_key = key
            enc = self.tokenize(tok, texts)
            out = {k: v.to(self.device, non_blocking=True) for k, v in enc.items()}
            out.update(self._image_fe(images))
            return out
        encodings = self.processor(images=images, text=texts, max_length=self.max_text_length,
                                   padding=True, truncation=True, return_tensors='pt').to(self.device)
        return encodings

    def expand_modality_type_embeddings(self, type_vocab_size=3):
        """vilt.py:98-109: third modality row (second NLVR2 image) initialised from row 1."""
        self.vilt.config.modality_type_vocab_size = type_vocab_size
        old = self.vilt.embeddings.token_type_embeddings.weight.data
        new = nn.Embedding(type_vocab_size, self.encoder_dim).to(old.device)
        new.weight.data[0, :] = old[0, :]
        new.weight.data[1, :] = old[1, :]
        new.weight.data[2, :] = old[1, :]
        self.vilt.embeddings.token_type_embeddings = new

    def forward(self, **encodings) -> torch.FloatTensor:
        return self.vilt(**encodings).pooler_output

    def freeze_all_weights(self):
        for p in self.vilt.parameters():
            p.requires_grad = False

    def freeze_bottom_k_layers(self, k: int):
        assert k < len(self.vilt.encoder.layer)
        for p in self.vilt.embeddings.parameters():
            p.requires_grad = False
        for i in range(k):
            for p in self.vilt.encoder.layer[i].parameters():
                p.requires_grad = False


class ClassifierHead(nn.Sequential):
    """nn.Sequential(Linear, LayerNorm, GELU, Linear) of vilt.py:190-197 -- same child indices, hence the
    same state-dict keys (task_layer.<task>.{0,1,3}.*) -- evaluated by the CUDA kernels."""

    def __init__(self, in_dim: int, hidden: int, num_labels: int):
        super().__init__(nn.Linear(in_dim, hidden), nn.LayerNorm(hidden), nn.GELU(), nn.Linear(hidden, num_labels))

    def forward(self, x):
        z = ops.linear(x, self[0].weight, self[0].bias)
        z = ops.layernorm_gelu(z, self[1].weight, self[1].bias, self[1].eps, gelu=True)
        return ops.linear(z, self[3].weight, self[3].bias)


class MultiChoiceHead(nn.Sequential):
    """nn.Sequential(Dropout(0.1), Linear(d, 1)) of vilt.py:199-202."""

    def __init__(self, in_dim: int):
        super().__init__(nn.Dropout(0.1), nn.Linear(in_dim, 1))

    def forward(self, x):
        x = self[0](x)
        return ops.linear(x, self[1].weight, self[1].bias)


class B200ViltContinualLearner(ContinualLearner):
    def __init__(self, ordered_cl_tasks: List[str], encoder: B200ViltEncoderWrapper, encoder_dim: int, task_configs: Dict):
        super().__init__()
        self.encoder_dim = encoder_dim
        self.vilt_encoder = encoder
        self.ordered_cl_tasks = ordered_cl_tasks
        self.task_configs = task_configs
        self.task_layer_dict = {}
        for task_key in ordered_cl_tasks:
            self.add_task_layer(task_key, task_configs[task_key])
        self.task_layer = nn.ModuleDict(self.task_layer_dict)
        if 'nlvr2' in ordered_cl_tasks:
            self.vilt_encoder.expand_modality_type_embeddings()

    def add_task_layer(self, task_key: str, task_config: Dict):
        num_labels = task_config['num_labels']
        if task_config['model_type'] == 'classification':
            num_images = task_config['num_images']
            self.task_layer_dict[task_key] = ClassifierHead(self.encoder_dim * num_images, self.encoder_dim * 2, num_labels)
        elif task_config['model_type'] == 'multi-choice':
            self.task_layer_dict[task_key] = MultiChoiceHead(self.encoder_dim)

    def create_optimizer(self, hparams):
        """vilt.py:205-215, including its grouping quirk: only names containing 'bias' or
        'LayerNorm.weight' skip weight decay. Returns a torch Optimizer whose
        step() is one fused kernel per parameter arena."""
        no_decay = ['bias', 'LayerNorm.weight']
        groups = [
            {'params': [p for n, p in self.named_parameters() if not any(nd in n for nd in no_decay)],
             'weight_decay': hparams['weight_decay']},
            {'params': [p for n, p in self.named_parameters() if any(nd in n for nd in no_decay)],
             'weight_decay': 0.0}]
        return ArenaAdamW(groups, lr=hparams['lr'], eps=hparams['adam_epsilon'], betas=(0.9, 0.98),
                          arenas=[self.vilt_encoder.vilt._arena])

    def train(self, mode: bool = True):
        """nn.Module.train / eval, plus: tell the optional rank-slicing batch converter (climb_b200.distributed.sharding) whether
        the harness is training (batches are cut to this rank's rows) or evaluating (replicated, train_vqa.py:246-282)."""
        from ..distributed import set_training_mode
        set_training_mode(mode)
        return super().train(mode)

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, task_key: str, images: List, texts: List[str]):
        if getattr(self, "_ddp_forward_shard", False) and self.training:
            # the unchanged harness under torchrun (climb_b200.distributed.attach_if_distributed): this rank's rows only,
            # outputs all-gathered to the whole batch
            from ..distributed import sharded_forward
            return sharded_forward(lambda im, tx: self._forward_rows(task_key, im, tx), images, texts)
        return self._forward_rows(task_key, images, texts)

    def _forward_rows(self, task_key: str, images: List, texts: List[str]):
        task_config = self.task_configs[task_key]
        if task_config['model_type'] == 'multi-choice':
            texts = list(itertools.chain(*texts))
        elif task_config.get('num_images', 1) > 1:
            images = list(itertools.chain(*images))
        encodings = self.vilt_encoder.process_inputs(images, texts)
        return self.forward_tensors(task_key, encodings)

    def forward_tensors(self, task_key: str, encodings: Dict):
        """encodings as ViltProcessor lays them out: NLVR2 pixel_values [bs*2, 3, H, W] (the two images of a
        sample adjacent), VCR input_ids [bs*4, T] (the four choices adjacent)."""
        task_config = self.task_configs[task_key]
        if task_config['model_type'] == 'multi-choice':
            return self.forward_multi_choice(task_key, encodings, task_config['num_choices'])
        if task_config['num_images'] == 1:
            return self.forward_single_image(task_key, encodings)
        return self.forward_multi_images(task_key, encodings, task_config['num_images'])

    @staticmethod
    def _enc(encodings, key, default=None):
        try:
            return encodings[key]
        except (KeyError, TypeError):
            return getattr(encodings, key, default)

    def forward_single_image(self, task_key, encodings):
        pooled = self.vilt_encoder(**{k: encodings[k] for k in ('input_ids', 'attention_mask', 'token_type_ids',
                                                                'pixel_values', 'pixel_mask') if k in encodings})
        return pooled, self.task_layer[task_key](pooled)

    def forward_multi_images(self, task_key, encodings, num_images=2):
        """vilt.py:263-307, batched: sequence (b, i) = text b with image i and image_token_type_idx i + 1."""
        ids, am, tt = encodings['input_ids'], encodings['attention_mask'], encodings['token_type_ids']
        bs = len(ids)
        px = encodings['pixel_values']
        rep = lambda t: t.repeat_interleave(num_images, dim=0)
        type_idx = (torch.arange(num_images, device=px.device, dtype=torch.int32) + 1).repeat(bs)
        pooled = self.vilt_encoder(input_ids=rep(ids), attention_mask=rep(am), token_type_ids=rep(tt),
                                   pixel_values=px, pixel_mask=self._enc(encodings, 'pixel_mask'),
                                   image_token_type_idx=type_idx,
                                   # the reference runs image 0 of every sample, then image 1 (vilt.py:291-304)
                                   patch_draw_order=[b * num_images + i for i in range(num_images) for b in range(bs)])
        pooled = pooled.view(bs, num_images * pooled.shape[-1])       # == torch.cat(pooler_outputs, dim=-1)
        return pooled, self.task_layer[task_key](pooled)

    def forward_multi_choice(self, task_key, encodings, num_choices):
        """vilt.py:309-350, batched: sequence (b, c) = image b with text choice c."""
        px = encodings['pixel_values']
        bs = px.shape[0]
        pm = self._enc(encodings, 'pixel_mask')
        # the reference feeds the same pixels once per choice (vilt.py:334-347); here every image is embedded once and its
        # patch rows are shared by its num_choices sequences (climb_vilt_batch.image_repeat)
        pooled = self.vilt_encoder(input_ids=encodings['input_ids'], attention_mask=encodings['attention_mask'],
                                   token_type_ids=encodings['token_type_ids'], pixel_values=px, pixel_mask=pm,
                                   image_repeat=num_choices,
                                   # the reference runs choice 0 of every sample, then choice 1 ... (vilt.py:334-347)
                                   patch_draw_order=[b * num_choices + c for c in range(num_choices) for b in range(bs)])
        pooled = pooled.view(bs, num_choices, -1)                     # == stack(dim=0).transpose(0, 1)
        logits = self.task_layer[task_key](pooled).squeeze()
        return pooled, logits

    def get_encoder(self):
        return self.vilt_encoder

    # ---- adapter passthrough (vilt.py:356-367) --------------------------------------------------
    def add_adapter(self, task_key: str, config: Dict):
        self.vilt_encoder.vilt.add_adapter(task_key, config)

    def train_adapter(self, task_key: str):
        self.vilt_encoder.vilt.train_adapter(task_key)

    def set_active_adapters(self, task_key: str):
        self.vilt_encoder.vilt.set_active_adapters(task_key)

    def get_active_adapters(self):
        return self.vilt_encoder.vilt.active_adapters


def _resolve_pretrained(pretrained_vilt_name, processor, config):
    """Processor and config as load_vilt_encoder resolves them (vilt.py:498, 505): ViltProcessor.from_pretrained /
    ViltConfig.from_pretrained of `pretrained_vilt_name` (a hub name or a local directory; stock or vendored
    transformers), unless the caller handed them in."""
    if isinstance(pretrained_vilt_name, str):
        if processor is None:
            from transformers import ViltProcessor
            processor = ViltProcessor.from_pretrained(pretrained_vilt_name)
        if config is None:
            from transformers import ViltConfig
            config = ViltConfig.from_pretrained(pretrained_vilt_name)
    elif config is None and pretrained_vilt_name is not None:
        config = pretrained_vilt_name             # a config object / dict stands in for the hub name (offline use)
    return processor, config


def _read_checkpoint(checkpoint_name, state_dict):
    import os
    if state_dict is not None:
        return dict(state_dict)
    if isinstance(checkpoint_name, str) and os.path.isfile(checkpoint_name):
        return torch.load(checkpoint_name, map_location="cpu")
    return None


def _load_vilt_weights(enc: "B200ViltEncoderWrapper", checkpoint_name, sd: Dict) -> Dict:
    """The pre-finetuned branch of load_vilt_encoder (vilt.py:504-512): a checkpoint written by
    `encoder.state_dict()` (keys `vilt.*`; a bare ViltModel state dict is accepted too). The modality table grows
    to three rows when the checkpoint was trained with NLVR2 -- the reference tests the file NAME, the
    tensor's own shape is checked as well. Returns the entries that do not belong to ViltModel (`bert.*`)."""
    type_key = "embeddings.token_type_embeddings.weight"
    vilt_sd, rest = {}, {}
    for k, v in sd.items():
        if k.startswith("vilt."):
            vilt_sd[k[len("vilt."):]] = v
        elif k.startswith("bert."):
            rest[k] = v
        else:
            vilt_sd[k] = v
    rows = vilt_sd[type_key].shape[0] if type_key in vilt_sd else 0
    have = enc.vilt.embeddings.token_type_embeddings.weight.shape[0]
    if have < 3 and (rows == 3 or (isinstance(checkpoint_name, str) and 'nlvr2' in checkpoint_name and rows in (0, 3))):
        enc.expand_modality_type_embeddings()
    missing, unexpected = enc.vilt.load_state_dict(vilt_sd, strict=False)
    if unexpected:
        raise RuntimeError(f"unexpected keys in ViLT checkpoint: {unexpected[:5]} ...")
    hard_missing = [m for m in missing if "position_ids" not in m and ".adapters." not in m]
    if hard_missing:
        raise RuntimeError(f"ViLT checkpoint misses {len(hard_missing)} tensors: {hard_missing[:5]} ...")
    return rest


def load_vilt_encoder(checkpoint_name, device, pretrained_vilt_name=None, *, processor=None, config=None,
                      state_dict=None) -> B200ViltEncoderWrapper:
    """load_vilt_encoder of vilt.py:481-514, same positional signature `(checkpoint_name, device,
    pretrained_vilt_name)` -- the callers pass all three positionally (train_language.py:279, train_vision.py:311).

      checkpoint_name == pretrained_vilt_name (a hub name / local directory): ViltModel.from_pretrained weights;
      otherwise `checkpoint_name` is a file written by `torch.save(encoder.state_dict())`: the model is built from
      the config of `pretrained_vilt_name`, the modality table is expanded for NLVR2 checkpoints, the weights are loaded.

    Offline extensions (keyword-only, plus: a config object / dict in place of either name means random init with
    that config and no processor lookup): `processor`, `config`, `state_dict`."""
    if pretrained_vilt_name is None:
        pretrained_vilt_name = checkpoint_name          # what create_vilt_continual_learner_model passes (vilt.py:536-538)
    logger.info("Loading ViLT encoder model: %s", checkpoint_name)
    sd = _read_checkpoint(checkpoint_name, state_dict)
    if sd is None and isinstance(checkpoint_name, str):
        if checkpoint_name != pretrained_vilt_name:
            raise FileNotFoundError(f"ViLT encoder checkpoint not found: {checkpoint_name}")
        from transformers import ViltModel               # the pretrained branch (vilt.py:500-502)
        hf = ViltModel.from_pretrained(pretrained_vilt_name)
        config = config if config is not None else hf.config
        sd = hf.state_dict()
    if not isinstance(checkpoint_name, str) and config is None:
        config = checkpoint_name
    processor, config = _resolve_pretrained(pretrained_vilt_name, processor, config)
    enc = B200ViltEncoderWrapper(processor, B200ViltModel(config), device)
    if sd is not None:
        rest = _load_vilt_weights(enc, checkpoint_name, sd)
        if rest:
            raise RuntimeError(f"unexpected keys in ViLT checkpoint: {sorted(rest)[:5]} ... (a ViLT-BERT checkpoint? "
                               "use load_viltbert_encoder)")
    enc.to(device)
    logger.info("Successfully loaded pretrained ViLT encoder")
    return enc


def create_vilt_continual_learner_model(model_name_or_path, ordered_cl_tasks, model_config, task_configs, device, *,
                                        processor=None):
    """create_vilt_continual_learner_model of vilt.py:516-546 (same positional signature). When torch.distributed is
    initialised with more than one rank the learner's gradients are averaged over the ranks after every backward
    (climb_b200.distributed.attach): with the rank-slicing batch converter of the registry this is all a torchrun
    launch of the unchanged harness needs."""
    encoder = load_vilt_encoder(model_name_or_path, device, model_name_or_path, processor=processor)
    model = B200ViltContinualLearner(ordered_cl_tasks=ordered_cl_tasks, encoder=encoder,
                                     encoder_dim=model_config['encoder_dim'], task_configs=task_configs)
    model.to(device)
    from ..distributed import attach_if_distributed
    attach_if_distributed(model)
    return model


def convert_batch_to_vilt_input_dict(batch: Dict):
    return {'images': batch['images'], 'texts': batch['raw_texts']}


def convert_seq_batch_to_vilt_input_dict(batch: List, mean_image):
    return {'images': [mean_image], 'texts': list(batch[0])}
