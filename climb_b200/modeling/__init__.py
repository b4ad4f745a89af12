"""Encoder registry with the reference's shape (src/modeling/__init__.py:4-12).

A CLiMB checkout gains the B200 encoder by merging these two dicts into its own maps and the
model_configs entry below into src/configs/model_configs.py (INTEGRATION.md shows the three-line patch).
"""
from .continual_learner import ContinualLearner, EncoderWrapper
from .vilt import (B200ViltContinualLearner, B200ViltEncoderWrapper, convert_batch_to_vilt_input_dict,
                   create_vilt_continual_learner_model, load_vilt_encoder)
from .vilt_model import AdapterSpec, B200ViltConfig, B200ViltModel
from .bert_model import B200BertConfig, B200BertModel
from .downstream import (B200ViltBertForMultipleChoice, B200ViltBertForSequenceClassification, B200ViltForImageClassification,
                         B200ViltForMultipleChoice, B200ViltForSequenceClassification)
from .viltbert import (B200ViltBertContinualLearner, B200ViltBertEncoderWrapper, convert_batch_to_viltbert_input_dict,
                       create_viltbert_continual_learner_model, load_viltbert_encoder)

load_encoder_map = {
    'vilt-b200': load_vilt_encoder,
    'viltbert-b200': load_viltbert_encoder,
}

create_continual_learner_map = {
    'vilt-b200': create_vilt_continual_learner_model,
    'viltbert-b200': create_viltbert_continual_learner_model,
}

# src/configs/model_configs.py:6-12
vilt_b200_config = {
    'encoder_dim': 768,
    'visual_input_type': 'pil-image',
    'encoder_class': B200ViltEncoderWrapper,
    'batch2inputs_converter': convert_batch_to_vilt_input_dict,     # (rank sharding under torchrun happens inside the learner's forward: distributed.py)
    'encoder_name': 'ViLT-B200',
}
# src/configs/model_configs.py:36-42
viltbert_b200_config = {
    'encoder_dim': 768,
    'visual_input_type': 'pil-image',
    'encoder_class': B200ViltBertEncoderWrapper,
    'batch2inputs_converter': convert_batch_to_viltbert_input_dict,
    'encoder_name': 'ViLT-BERT-B200',
}
model_configs = {'vilt-b200': vilt_b200_config, 'viltbert-b200': viltbert_b200_config}
ALLOWED_CL_ENCODERS = ['vilt-b200', 'viltbert-b200']
