"""Item-table build (ref: finetune.py:38-63 `encode_all_items`, evaluate_seq.py:17-30).

Same contract as the reference helper — items sorted by id, each encoded as a one-item sequence
`<s> + item tokens` through `tokenizer.batch_encode(..., encode_item=False)`, CLS pooled — with two
B200-side changes: batches are assembled on the GPU from CSR token arrays (`item_store`) or in pinned host memory
copied asynchronously, only the CLS rows leave the encoder, and the pooled vectors can be written straight into the
L2-normalised bf16 shard that full-catalogue scoring reads (`normalized_out`), so a rank that owns item ids [lo, hi)
never materialises the fp32 table of the others.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops


@torch.no_grad()
def encode_all_items(model, tokenizer, tokenized_items: Dict[int, list], batch_size: int = 512, device=None,
                     id_range: Optional[tuple] = None, normalized_out: Optional[torch.Tensor] = None,
                     item_store=None) -> torch.Tensor:
    """Returns the fp32 [n_items, E] table of CLS vectors (rows in ascending item id, ref: finetune.py:42-43).

    model: RecformerModel (or anything with `.longformer`); id_range=(lo, hi) restricts to the ids a rank owns
    (sharded scoring); normalized_out: optional bf16 [n_items, E] buffer that receives the
    L2-normalised rows (what `rf_cosine_topk` consumes).

    Batch assembly: `item_store` (a `tokenization.DeviceItemStore`, built once per catalogue) assembles each batch's
    five [B, L] tensors ON THE GPU from the CSR token arrays (`rf_assemble_batch`, bit-identical to
    `tokenizer.batch_encode(..., encode_item=False)`), so an epoch's re-encoding of the catalogue
    (ref: finetune.py:304-307) no longer runs the tokenizer's Python loops per batch; `item_store=True` builds the store
    here.  Without it the reference's host path is used (pinned buffers, asynchronous copies).  Only the CLS rows leave
    the encoder (`RecformerModel.forward_pooled`): no [B, L, E] fp32 hidden-state copy per batch."""
    enc = getattr(model, "longformer", model)
    device = torch.device(device) if device is not None else next(enc.parameters()).device
    was_training = enc.training
    enc.eval()
    ids = sorted(tokenized_items)
    if id_range is not None:
        ids = [i for i in ids if id_range[0] <= i < id_range[1]]
    if item_store is True:
        from .tokenization import DeviceItemStore
        item_store = DeviceItemStore(enc.config, {i: tokenized_items[i] for i in ids}, device=device)
    E = enc.config.hidden_size
    table = torch.empty(len(ids), E, dtype=torch.float32, device=device)
    for a in range(0, len(ids), batch_size):
        chunk = ids[a:a + batch_size]
        if item_store is not None:
            dev = item_store.batch_encode([[i] for i in chunk])
        else:
            inputs = tokenizer.batch_encode([[tokenized_items[i]] for i in chunk], encode_item=False)
            dev = {k: torch.tensor(v, dtype=torch.int64).pin_memory().to(device, non_blocking=True) for k, v in inputs.items()}
        pooled = enc.forward_pooled(**dev)
        table[a:a + len(chunk)] = pooled
        if normalized_out is not None:
            ops.normalize_rows(pooled.contiguous(), out=normalized_out[a:a + len(chunk)])
    if was_training:
        enc.train()
    return table
