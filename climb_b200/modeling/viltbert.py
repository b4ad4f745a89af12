"""B200 drop-ins for CLiMB's ViLT-BERT wrappers (src/modeling/viltbert.py): same class surface, registry
signatures and state-dict keys (`viltbert_encoder.vilt.*`, `viltbert_encoder.bert.*`, `task_layer.*`).

    ViltBertEncoderWrapper   -> B200ViltBertEncoderWrapper   (viltbert.py:31-168)
    ViltBertContinualLearner -> B200ViltBertContinualLearner (viltbert.py:171-371)
    load_viltbert_encoder / create_viltbert_continual_learner_model (viltbert.py:456-520)

The frozen BERT runs first (climb_bert_forward, forward only, as under the reference's no_grad), its
last_hidden_state enters the ViLT engine as inputs_embeds, so `word_embeddings.weight` of the ViLT text
embeddings gets no gradient. Deliberate differences:
  * NLVR2's two and VCR's four encoder passes are one batched call; for NLVR2 BERT runs ONCE per text
    (the reference recomputes the identical features for the second image, viltbert.py:291-304);
  * the learner has create_optimizer / get_active_adapters, which the reference class lacks although the
    trainers and EWC call them -- same semantics as vilt.py:205-215, 366-367.
"""
from __future__ import annotations

import logging
from typing import Dict, List

import torch
import torch.nn as nn

from ..optim import ArenaAdamW
from .bert_model import B200BertConfig, B200BertModel
from .vilt import B200ViltContinualLearner, B200ViltEncoderWrapper
from .vilt_model import B200ViltModel

logger = logging.getLogger(__name__)


class B200ViltBertEncoderWrapper(B200ViltEncoderWrapper):
    def __init__(self, processor, vilt: B200ViltModel, bert: B200BertModel, device: torch.device):
        super().__init__(processor, vilt, device)
        self.bert = bert
        if bert.config.hidden_size != vilt.config.hidden_size:
            raise ValueError("BERT and ViLT hidden sizes differ: BERT's last_hidden_state is ViLT's inputs_embeds")

    def get_bert_outputs(self, **encodings) -> torch.Tensor:
        """viltbert.py:115-120 (no_grad)."""
        return self.bert(input_ids=encodings['input_ids'], attention_mask=encodings['attention_mask'],
                         token_type_ids=encodings['token_type_ids']).last_hidden_state

    def create_optimizer(self, hparams):
        """viltbert.py:122-132 (the reference defines it on the wrapper)."""
        no_decay = ['bias', 'LayerNorm.weight']
        groups = [
            {'params': [p for n, p in self.named_parameters() if not any(nd in n for nd in no_decay)],
             'weight_decay': hparams['weight_decay']},
            {'params': [p for n, p in self.named_parameters() if any(nd in n for nd in no_decay)], 'weight_decay': 0.0}]
        return ArenaAdamW(groups, lr=hparams['lr'], eps=hparams['adam_epsilon'], betas=(0.9, 0.98), arenas=[self.vilt._arena])

    def forward(self, **encodings) -> torch.FloatTensor:
        """viltbert.py:135-151. `inputs_embeds` may be passed in (features computed once for several image passes)."""
        enc = dict(encodings)
        if enc.get('inputs_embeds') is None:
            enc['inputs_embeds'] = self.get_bert_outputs(**enc)
        enc['input_ids'] = None
        return self.vilt(**enc).pooler_output


class B200ViltBertContinualLearner(B200ViltContinualLearner):
    """Same forward logic as the ViLT learner over a ViLT-BERT encoder; the attribute is named
    `viltbert_encoder` as in the reference so that checkpoints keep their keys."""

    def __init__(self, ordered_cl_tasks: List[str], encoder: B200ViltBertEncoderWrapper, encoder_dim: int, task_configs: Dict):
        nn.Module.__init__(self)
        self.encoder_dim = encoder_dim
        self.viltbert_encoder = encoder
        self.ordered_cl_tasks = ordered_cl_tasks
        self.task_configs = task_configs
        self.task_layer_dict = {}
        for task_key in ordered_cl_tasks:
            self.add_task_layer(task_key, task_configs[task_key])
        self.task_layer = nn.ModuleDict(self.task_layer_dict)
        if 'nlvr2' in ordered_cl_tasks:
            self.viltbert_encoder.expand_modality_type_embeddings()

    @property
    def vilt_encoder(self):          # the shared forward code of B200ViltContinualLearner reads this name
        return self.viltbert_encoder

    def forward_multi_images(self, task_key, encodings, num_images=2):
        """viltbert.py:275-318, batched; BERT features computed once per text and repeated per image."""
        ids, am, tt = encodings['input_ids'], encodings['attention_mask'], encodings['token_type_ids']
        bs = len(ids)
        px = encodings['pixel_values']
        feats = self.viltbert_encoder.get_bert_outputs(input_ids=ids, attention_mask=am, token_type_ids=tt)
        rep = lambda t: t.repeat_interleave(num_images, dim=0)
        type_idx = (torch.arange(num_images, device=px.device, dtype=torch.int32) + 1).repeat(bs)
        pooled = self.viltbert_encoder(input_ids=None, inputs_embeds=rep(feats), attention_mask=rep(am),
                                       token_type_ids=rep(tt), pixel_values=px,
                                       pixel_mask=self._enc(encodings, 'pixel_mask'),
                                       image_token_type_idx=type_idx,
                                       patch_draw_order=[b * num_images + i for i in range(num_images) for b in range(bs)])
        pooled = pooled.view(bs, num_images * pooled.shape[-1])
        return pooled, self.task_layer[task_key](pooled)

    def get_encoder(self):
        return self.viltbert_encoder


def load_viltbert_encoder(checkpoint_name, device, pretrained_vilt_name=None, *, processor=None, config=None,
                          state_dict=None, bert=None, bert_config=None) -> B200ViltBertEncoderWrapper:
    """load_viltbert_encoder of viltbert.py:456-489, same positional signature `(checkpoint_name, device,
    pretrained_vilt_name)`. A checkpoint file holds `vilt.*` and `bert.*` keys (the wrapper's own state dict).
    Keyword-only offline extensions: `bert` may be a B200BertModel or a BERT state dict, `bert_config` asks for a
    random-init BERT; otherwise BertModel.from_pretrained('bert-base-uncased') as in the reference (:474)."""
    from .vilt import _load_vilt_weights, _read_checkpoint, _resolve_pretrained
    if pretrained_vilt_name is None:
        pretrained_vilt_name = checkpoint_name
    logger.info("Loading ViLT encoder model: %s", checkpoint_name)
    sd = _read_checkpoint(checkpoint_name, state_dict)
    if sd is None and isinstance(checkpoint_name, str):
        if checkpoint_name != pretrained_vilt_name:
            raise FileNotFoundError(f"ViLT-BERT encoder checkpoint not found: {checkpoint_name}")
        from transformers import ViltModel
        hf = ViltModel.from_pretrained(pretrained_vilt_name)
        config = config if config is not None else hf.config
        sd = hf.state_dict()
    if not isinstance(checkpoint_name, str) and config is None:
        config = checkpoint_name
    processor, config = _resolve_pretrained(pretrained_vilt_name, processor, config)
    ckpt_bert = {k[len("bert."):]: v for k, v in (sd or {}).items() if k.startswith("bert.")}
    if not isinstance(bert, B200BertModel):
        bert_sd = bert if bert is not None else (ckpt_bert or None)
        if bert_config is None:
            if bert_sd is None:
                from transformers import BertModel
                hf = BertModel.from_pretrained("bert-base-uncased")
                bert_config, bert_sd = hf.config, hf.state_dict()
            else:
                bert_config = B200BertConfig()           # bert-base-uncased geometry
        bert = B200BertModel(bert_config)
        if bert_sd is not None:
            missing, unexpected = bert.load_state_dict(bert_sd, strict=False)
            if unexpected:
                raise RuntimeError(f"unexpected keys in BERT checkpoint: {unexpected[:5]} ...")
    elif ckpt_bert:
        bert.load_state_dict(ckpt_bert, strict=False)
    enc = B200ViltBertEncoderWrapper(processor, B200ViltModel(config), bert, device)
    if sd is not None:
        _load_vilt_weights(enc, checkpoint_name, sd)
    enc.to(device)
    logger.info("Successfully loaded pretrained ViLT-BERT encoder")
    return enc


def create_viltbert_continual_learner_model(model_name_or_path, ordered_cl_tasks, model_config, task_configs, device, *,
                                            processor=None, bert=None, bert_config=None):
    """create_viltbert_continual_learner_model of viltbert.py:491-520 (same positional signature)."""
    encoder = load_viltbert_encoder(model_name_or_path, device, model_name_or_path, processor=processor, bert=bert,
                                    bert_config=bert_config)
    model = B200ViltBertContinualLearner(ordered_cl_tasks=ordered_cl_tasks, encoder=encoder,
                                         encoder_dim=model_config['encoder_dim'], task_configs=task_configs)
    model.to(device)
    from ..distributed import attach_if_distributed
    attach_if_distributed(model)
    return model


def convert_batch_to_viltbert_input_dict(batch: Dict):
    return {'images': batch['images'], 'texts': batch['raw_texts']}
