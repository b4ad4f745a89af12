"""Downstream classifiers that reuse the upstream-trained encoder (src/modeling/vilt.py:370-478 and the
ViLT-BERT twins src/modeling/viltbert.py:374-453): vision-only (dummy text), language-only sequence
classification and multiple choice (one "mean image" broadcast over the batch). Same attribute names
(`vilt_encoder` / `encoder`, `clf_layer`) and state-dict keys as the reference classes; the encoder is
B200ViltEncoderWrapper or B200ViltBertEncoderWrapper, the heads run on the CUDA kernels of climb_b200.ops."""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from .vilt import ClassifierHead, MultiChoiceHead


def _broadcast_image(encodings: Dict, bs: int) -> Dict:
    """The language-only tasks pass ONE image for the whole batch (vilt.py:439-441, 471-473)."""
    enc = dict(encodings)
    px = enc['pixel_values']
    enc['pixel_values'] = px.expand([bs, *px.shape[1:]])
    pm = enc.get('pixel_mask')
    if pm is not None:
        enc['pixel_mask'] = pm.expand([bs, *pm.shape[1:]])
    return enc


class B200ViltForImageClassification(nn.Module):
    """ViltForImageClassification, vilt.py:370-403."""

    def __init__(self, encoder, encoder_dim: int, num_labels: int):
        super().__init__()
        self.encoder_dim = encoder_dim
        self.vilt_encoder = encoder
        self.clf_layer = ClassifierHead(encoder_dim, encoder_dim * 2, num_labels)

    def forward(self, images: List, texts: List[str]) -> torch.FloatTensor:
        return self.forward_tensors(self.vilt_encoder.process_inputs(images, texts))

    def forward_tensors(self, encodings: Dict) -> torch.FloatTensor:
        return self.clf_layer(self.vilt_encoder(**encodings))


class B200ViltForSequenceClassification(nn.Module):
    """ViltForSequenceClassification, vilt.py:406-443."""

    def __init__(self, encoder, encoder_dim: int, num_labels: int):
        super().__init__()
        self.encoder_dim = encoder_dim
        self.encoder = encoder
        self.clf_layer = ClassifierHead(encoder_dim, encoder_dim * 2, num_labels)

    def forward(self, images: List, texts: List[str]) -> torch.FloatTensor:
        return self.forward_tensors(self.encoder.process_inputs(images, texts))

    def forward_tensors(self, encodings: Dict) -> torch.FloatTensor:
        enc = _broadcast_image(encodings, len(encodings['input_ids']))
        return self.clf_layer(self.encoder(**enc))


class B200ViltForMultipleChoice(nn.Module):
    """ViltForMultipleChoice, vilt.py:446-478: texts arrive choice-major ([num_labels * bs] rows), logits [bs, num_labels]."""

    def __init__(self, encoder, encoder_dim: int, num_labels: int):
        super().__init__()
        self.encoder_dim = encoder_dim
        self.num_labels = num_labels
        self.encoder = encoder
        self.clf_layer = MultiChoiceHead(encoder_dim)

    def forward(self, images, texts):
        return self.forward_tensors(self.encoder.process_inputs(images, texts))

    def forward_tensors(self, encodings: Dict) -> torch.FloatTensor:
        enc = _broadcast_image(encodings, len(encodings['input_ids']))
        out = self.encoder(**enc)
        out = out.view(self.num_labels, -1, self.encoder_dim).transpose(0, 1).contiguous()
        return self.clf_layer(out).squeeze()


# the ViLT-BERT variants of the reference are the same classes over a ViLT-BERT encoder wrapper
B200ViltBertForSequenceClassification = B200ViltForSequenceClassification
B200ViltBertForMultipleChoice = B200ViltForMultipleChoice
