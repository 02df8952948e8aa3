"""RecformerConfig — same field names and defaults as the reference (ref: recformer/models.py:24-55)
on top of the public longformer-base-4096 values (the reference builds it with
`RecformerConfig.from_pretrained('allenai/longformer-base-4096')`, ref: finetune.py:203, which
needs the hub; here the literals are built in)."""
from __future__ import annotations

import copy
import json
import os
from typing import List, Union

_LONGFORMER_BASE = dict(
    vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
    hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=4098,
    type_vocab_size=1, initializer_range=0.02, layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2,
    attention_window=[512] * 12,
)


class RecformerConfig:
    model_type = "longformer"

    def __init__(self,
                 attention_window: Union[List[int], int] = 64,
                 sep_token_id: int = 2,
                 token_type_size: int = 4,        # <s>, key, value, <pad>
                 max_token_num: int = 2048,
                 max_item_embeddings: int = 32,   # 1 for <s>, 50 for items
                 max_attr_num: int = 12,
                 max_attr_length: int = 8,
                 pooler_type: str = "cls",
                 temp: float = 0.05,
                 mlm_weight: float = 0.1,
                 item_num: int = 0,
                 finetune_negative_sample_size: int = 0,
                 **kwargs):
        base = dict(_LONGFORMER_BASE)
        base.update(kwargs)
        for k, v in base.items():
            setattr(self, k, v)
        self.attention_window = attention_window
        self.sep_token_id = sep_token_id
        self.token_type_size = token_type_size
        self.max_token_num = max_token_num
        self.max_item_embeddings = max_item_embeddings
        self.max_attr_num = max_attr_num
        self.max_attr_length = max_attr_length
        self.pooler_type = pooler_type
        self.temp = temp
        self.mlm_weight = mlm_weight
        self.item_num = item_num
        self.finetune_negative_sample_size = finetune_negative_sample_size
        self.output_attentions = base.get("output_attentions", False)
        self.output_hidden_states = base.get("output_hidden_states", False)
        self.use_return_dict = base.get("return_dict", True)

    # -- the slice of the HF PretrainedConfig surface the reference scripts touch --------------
    @classmethod
    def from_pretrained(cls, name_or_path: str, **kwargs):
        """'allenai/longformer-base-4096' (or any *longformer-base* id) resolves to the built-in
        literals; a directory (or file) holding config.json is read from disk."""
        path = name_or_path
        if os.path.isdir(path):
            path = os.path.join(path, "config.json")
        if os.path.isfile(path):
            with open(path) as f:
                d = json.load(f)
            d.update(kwargs)
            aw = d.pop("attention_window", 64)
            return cls(attention_window=aw, **{k: v for k, v in d.items() if k not in ("model_type", "architectures")})
        if "longformer-base" in name_or_path:
            return cls(attention_window=list(_LONGFORMER_BASE["attention_window"]), **kwargs)
        raise OSError(f"RecformerConfig.from_pretrained: cannot resolve {name_or_path!r} offline")

    def to_dict(self):
        return copy.deepcopy({k: v for k, v in self.__dict__.items() if not k.startswith("_")})

    def save_pretrained(self, directory: str):
        os.makedirs(directory, exist_ok=True)
        with open(os.path.join(directory, "config.json"), "w") as f:
            json.dump(dict(self.to_dict(), model_type=self.model_type), f, indent=2)

    def __repr__(self):
        return f"RecformerConfig({json.dumps(self.to_dict(), default=str)})"
