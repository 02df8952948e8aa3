"""RecformerTokenizer batch layout (ref: recformer/tokenization.py:4-159; SURVEY.md §8a Spec T).

The reference subclasses HF's LongformerTokenizer; its BPE vocabulary comes from the hub and is
out of scope (no network).  The layout logic — `encode(items, encode_item=False)`, `padding`,
`batch_encode`, `__call__` — is reproduced exactly on pre-tokenised items
(`tokenized_items = {item_id: [input_ids, token_type_ids]}`, ref: finetune.py:239).  Attribute
text can still be tokenised by passing any callable `text_tokenizer(str) -> List[int]`."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch


class RecformerTokenizer:
    def __init__(self, config, text_tokenizer: Optional[Callable[[str], List[int]]] = None, bos_token_id: int = 0,
                 pad_token_id: int = 1):
        self.config = config
        self.text_tokenizer = text_tokenizer
        self.bos_token_id = bos_token_id
        self.pad_token_id = pad_token_id

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, config=None, text_tokenizer=None):
        return cls(config, text_tokenizer=text_tokenizer)

    def __call__(self, items, pad_to_max=False, return_tensor=False):
        if len(items) > 0 and isinstance(items[0], list):
            inputs = self.batch_encode(items, pad_to_max=pad_to_max)
        else:
            inputs = self.encode(items)
        if return_tensor:
            for k, v in inputs.items():
                inputs[k] = torch.LongTensor(v)
        return inputs

    def item_tokenize(self, text: str) -> List[int]:
        if self.text_tokenizer is None:
            raise RuntimeError("RecformerTokenizer: no text tokenizer configured (the BPE vocabulary is not "
                               "available offline); pass pre-tokenised items with encode_item=False")
        return list(self.text_tokenizer(text))

    def encode_item(self, item: Dict[str, str]):
        """ref: tokenization.py:38-61."""
        input_ids, token_type_ids = [], []
        for attr_name, attr_value in list(item.items())[: self.config.max_attr_num]:
            name_tokens = self.item_tokenize(attr_name)
            value_tokens = self.item_tokenize(attr_value)
            attr_tokens = (name_tokens + value_tokens)[: self.config.max_attr_length]
            input_ids += attr_tokens
            token_type_ids += ([1] * len(name_tokens) + [2] * len(value_tokens))[: self.config.max_attr_length]
        return input_ids, token_type_ids

    def encode(self, items, encode_item=True):
        """ref: tokenization.py:64-107 — [past..present] in, <s> + [present..past] out."""
        items = items[::-1]
        items = items[: self.config.max_item_embeddings - 1]
        input_ids, item_position_ids, token_type_ids = [self.bos_token_id], [0], [0]
        for item_idx, item in enumerate(items):
            item_input_ids, item_token_type_ids = self.encode_item(item) if encode_item else item
            input_ids += item_input_ids
            token_type_ids += item_token_type_ids
            item_position_ids += [item_idx + 1] * len(item_input_ids)
        n = self.config.max_token_num
        input_ids, item_position_ids, token_type_ids = input_ids[:n], item_position_ids[:n], token_type_ids[:n]
        attention_mask = [1] * len(input_ids)
        global_attention_mask = [0] * len(input_ids)
        global_attention_mask[0] = 1
        return {"input_ids": input_ids, "item_position_ids": item_position_ids, "token_type_ids": token_type_ids,
                "attention_mask": attention_mask, "global_attention_mask": global_attention_mask}

    def padding(self, item_batch, pad_to_max):
        """ref: tokenization.py:109-152."""
        max_length = self.config.max_token_num if pad_to_max else max(len(x["input_ids"]) for x in item_batch)
        fill = {"input_ids": self.pad_token_id, "item_position_ids": self.config.max_item_embeddings - 1,
                "token_type_ids": 3, "attention_mask": 0, "global_attention_mask": 0}
        out = {k: [] for k in fill}
        for x in item_batch:
            n = max_length - len(x["input_ids"])
            for k in out:
                out[k].append(list(x[k]) + [fill[k]] * n)
        return out

    def batch_encode(self, item_batch, encode_item=True, pad_to_max=False):
        return self.padding([self.encode(items, encode_item) for items in item_batch], pad_to_max)


class DeviceItemStore:
    """Pre-tokenised items (`{item_id: [input_ids, token_type_ids]}`, ref: finetune.py:239) held on the GPU as CSR
    arrays; `batch_encode` assembles the tokenizer's five-tensor batch there with one kernel (rf_assemble_batch)
    instead of Python list loops + a host->device copy of 5 x B x L int64.  Bit-identical to
    `RecformerTokenizer.batch_encode(..., encode_item=False)`."""

    def __init__(self, config, tokenized_items, device="cuda", bos_token_id: int = 0, pad_token_id: int = 1):
        import numpy as np
        self.config, self.device = config, torch.device(device)
        self.bos_token_id, self.pad_token_id = bos_token_id, pad_token_id
        self.ids = sorted(tokenized_items)
        self.index = {item_id: k for k, item_id in enumerate(self.ids)}
        lens = np.array([len(tokenized_items[i][0]) for i in self.ids], dtype=np.int64)
        offsets = np.zeros(len(self.ids) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        tokens = np.concatenate([np.asarray(tokenized_items[i][0], dtype=np.int32) for i in self.ids]) if len(self.ids) else np.zeros(0, np.int32)
        types = np.concatenate([np.asarray(tokenized_items[i][1], dtype=np.uint8) for i in self.ids]) if len(self.ids) else np.zeros(0, np.uint8)
        self.offsets = torch.from_numpy(offsets).to(self.device)
        self.tokens = torch.from_numpy(tokens).to(self.device)
        self.types = torch.from_numpy(types).to(self.device)

    def batch_encode(self, user_item_ids, pad_to_max: bool = False):
        """user_item_ids: list (per user) of item ids, oldest first (what `RecformerTokenizer.encode` receives)."""
        from . import _lib
        from .ops import check, _stream
        B = len(user_item_ids)
        flat = [self.index[i] for items in user_item_ids for i in items]
        uoff = [0]
        for items in user_item_ids:
            uoff.append(uoff[-1] + len(items))
        u_items = torch.tensor(flat if flat else [0], dtype=torch.int64).pin_memory().to(self.device, non_blocking=True)
        u_off = torch.tensor(uoff, dtype=torch.int64).pin_memory().to(self.device, non_blocking=True)
        cfg = self.config
        max_items, max_tokens = cfg.max_item_embeddings - 1, cfg.max_token_num
        if pad_to_max:
            L = max_tokens
        else:   # the batch maximum is known on the host from the item lengths (no device sync)
            lens_host = self._host_lengths(user_item_ids, max_items, max_tokens)
            L = max(lens_host)
        out = {k: torch.empty(B, L, dtype=torch.int64, device=self.device)
               for k in ("input_ids", "item_position_ids", "token_type_ids", "attention_mask", "global_attention_mask")}
        check(_lib.lib().rf_assemble_batch(self.offsets.data_ptr(), self.tokens.data_ptr(), self.types.data_ptr(),
                                           u_off.data_ptr(), u_items.data_ptr(), B, L, max_items, max_tokens,
                                           self.bos_token_id, self.pad_token_id, cfg.max_item_embeddings - 1,
                                           out["input_ids"].data_ptr(), out["item_position_ids"].data_ptr(),
                                           out["token_type_ids"].data_ptr(), out["attention_mask"].data_ptr(),
                                           out["global_attention_mask"].data_ptr(), None, _stream()), "rf_assemble_batch")
        return out

    def _host_lengths(self, user_item_ids, max_items, max_tokens):
        if not hasattr(self, "_len_host"):
            off = self.offsets.cpu()
            self._len_host = (off[1:] - off[:-1]).tolist()
        out = []
        for items in user_item_ids:
            n = 1 + sum(self._len_host[self.index[i]] for i in items[::-1][:max_items])
            out.append(min(n, max_tokens))
        return out
