"""Import the UNMODIFIED reference ``recformer/models.py`` (and utils / tokenization) by path.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the GPU box, so
nothing that runs there may import this module.  It exists to (a) pin the oracle against the
real reference and (b) generate the fixtures under tests/golden/.

The reference targets transformers 4.28; this image has 5.5.0.  Three API breaks are patched
*outside* the reference; reference files are never edited.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import torch

REF_ROOT = os.environ.get("RECFORMER_REF_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "recformer", "models.py"))


_cached = {}


def load_reference():
    """Returns the reference `recformer.models` module (classes RecformerConfig, RecformerModel,
    RecformerForSeqRec, ...)."""
    if "models" in _cached:
        return _cached["models"]
    from transformers import PreTrainedModel
    from transformers.models.longformer import modeling_longformer as ml
    from transformers.models.longformer.configuration_longformer import LongformerConfig

    if not getattr(ml, "_rf_shimmed", False):
        _enc = ml.LongformerEncoder.forward            # 5.5.0 dropped head_mask (ref passes None)

        def enc(self, hidden_states, attention_mask=None, head_mask=None, padding_len=0,
                output_attentions=False, output_hidden_states=False, return_dict=True):
            assert head_mask is None
            return _enc(self, hidden_states, attention_mask=attention_mask, padding_len=padding_len,
                        output_attentions=bool(output_attentions),
                        output_hidden_states=bool(output_hidden_states), return_dict=return_dict)

        ml.LongformerEncoder.forward = enc
        _gem = PreTrainedModel.get_extended_attention_mask   # 3rd positional is `dtype` now, ref passes device

        def gem(self, attention_mask, input_shape, device_or_dtype=None, dtype=None):
            return _gem(self, attention_mask, input_shape,
                        device_or_dtype if isinstance(device_or_dtype, torch.dtype) else dtype)

        ml.LongformerPreTrainedModel.get_extended_attention_mask = gem
        _ci = LongformerConfig.__init__                 # keyword-only config; ref calls it positionally

        def ci(self, *a, **kw):
            if len(a) > 0:
                kw["attention_window"] = a[0]
            if len(a) > 1:
                kw["sep_token_id"] = a[1]
            _ci(self, **kw)

        LongformerConfig.__init__ = ci
        ml._rf_shimmed = True

    spec = importlib.util.spec_from_file_location("ref_models", os.path.join(REF_ROOT, "recformer", "models.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_models"] = mod
    spec.loader.exec_module(mod)
    _cached["models"] = mod
    return mod


def load_reference_utils():
    if "utils" in _cached:
        return _cached["utils"]
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF_ROOT, "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_utils"] = mod
    spec.loader.exec_module(mod)
    _cached["utils"] = mod
    return mod


def load_reference_tokenization():
    if "tok" in _cached:
        return _cached["tok"]
    spec = importlib.util.spec_from_file_location("ref_tokenization",
                                                  os.path.join(REF_ROOT, "recformer", "tokenization.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_tokenization"] = mod
    spec.loader.exec_module(mod)
    _cached["tok"] = mod
    return mod


def reference_config(ocfg):
    """Build the reference RecformerConfig from an OracleConfig (literal values, no hub)."""
    ref = load_reference()
    cfg = ref.RecformerConfig(
        attention_window=list(ocfg.attention_window), vocab_size=ocfg.vocab_size, hidden_size=ocfg.hidden_size,
        num_hidden_layers=ocfg.num_hidden_layers, num_attention_heads=ocfg.num_attention_heads,
        intermediate_size=ocfg.intermediate_size, hidden_act="gelu", hidden_dropout_prob=0.1,
        attention_probs_dropout_prob=0.1, max_position_embeddings=ocfg.max_position_embeddings,
        type_vocab_size=1, layer_norm_eps=ocfg.layer_norm_eps, initializer_range=0.02,
        pad_token_id=ocfg.pad_token_id, bos_token_id=ocfg.bos_token_id, eos_token_id=2, sep_token_id=2)
    cfg.max_attr_num = ocfg.max_attr_num
    cfg.max_attr_length = ocfg.max_attr_length
    cfg.max_item_embeddings = ocfg.max_item_embeddings
    cfg.attention_window = list(ocfg.attention_window)
    cfg.max_token_num = ocfg.max_token_num
    cfg.token_type_size = ocfg.token_type_size
    cfg.temp = ocfg.temp
    cfg.pooler_type = ocfg.pooler_type
    cfg.item_num = ocfg.item_num
    cfg.finetune_negative_sample_size = ocfg.finetune_negative_sample_size
    return cfg
