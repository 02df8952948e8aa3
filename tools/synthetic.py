"""Seeded synthetic weights and inputs of the Recformer hot path: neutral ground.

Shared by the goldens (``tests/golden/make_goldens.py``), the tests, ``__graft_entry__.smoke()``, ``bench.py`` (both
arms) and the CPU oracle, so that every party regenerates bit-identical tensors from a seed on whichever machine it
runs (numpy's PCG64 stream is platform independent; no file travels).  Nothing here restates reference ARITHMETIC:
only shapes, key names and the tokenizer's 5-tensor batch layout.  The product package ``recformer_b200`` does not
import this module, and this module imports neither the product nor the oracle.

``SynthConfig`` carries the literal longformer-base / Recformer shape values (ref: recformer/models.py:24-55,
finetune.py:203-209; SURVEY.md §8d: the hub is unreachable, nothing is read from a checkpoint).  The generators take
any object with those attributes (``SynthConfig``, the oracle's ``OracleConfig`` alias, ``RecformerConfig``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------
# config (ref: recformer/models.py:24-55, finetune.py:203-209; literal longformer-base values,
# SURVEY.md §8d — the hub is unreachable so nothing is read from a checkpoint)
# ----------------------------------------------------------------------------------------------
@dataclass
class SynthConfig:
    vocab_size: int = 50265
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 4098
    layer_norm_eps: float = 1e-5
    pad_token_id: int = 1
    bos_token_id: int = 0
    attention_window: Sequence[int] = field(default_factory=lambda: [64] * 12)
    token_type_size: int = 4
    max_token_num: int = 1024
    max_item_embeddings: int = 51
    max_attr_num: int = 3
    max_attr_length: int = 32
    pooler_type: str = "cls"
    temp: float = 0.05
    item_num: int = 0
    finetune_negative_sample_size: int = 0

    def __post_init__(self):
        if isinstance(self.attention_window, int):
            self.attention_window = [self.attention_window] * self.num_hidden_layers


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs (shared by goldens, tests, smoke and bench)
# ----------------------------------------------------------------------------------------------
def state_dict_keys(cfg: SynthConfig, prefix: str = "") -> List[tuple]:
    """(key, shape, kind) for every parameter of RecformerModel, in the reference's key layout
    (ref: recformer/models.py:82-106,189-191; HF:445-1244)."""
    E, Fd = cfg.hidden_size, cfg.intermediate_size
    out = [
        (prefix + "embeddings.word_embeddings.weight", (cfg.vocab_size, E), "emb_pad"),
        (prefix + "embeddings.position_embeddings.weight", (cfg.max_position_embeddings, E), "emb_pad"),
        (prefix + "embeddings.token_type_embeddings.weight", (cfg.token_type_size, E), "emb"),
        (prefix + "embeddings.item_position_embeddings.weight", (cfg.max_item_embeddings, E), "emb"),
        (prefix + "embeddings.LayerNorm.weight", (E,), "ln_w"),
        (prefix + "embeddings.LayerNorm.bias", (E,), "ln_b"),
    ]
    for i in range(cfg.num_hidden_layers):
        p = f"{prefix}encoder.layer.{i}."
        for n in ("query", "key", "value", "query_global", "key_global", "value_global"):
            out.append((p + f"attention.self.{n}.weight", (E, E), "w"))
            out.append((p + f"attention.self.{n}.bias", (E,), "b"))
        out += [
            (p + "attention.output.dense.weight", (E, E), "w"),
            (p + "attention.output.dense.bias", (E,), "b"),
            (p + "attention.output.LayerNorm.weight", (E,), "ln_w"),
            (p + "attention.output.LayerNorm.bias", (E,), "ln_b"),
            (p + "intermediate.dense.weight", (Fd, E), "w"),
            (p + "intermediate.dense.bias", (Fd,), "b"),
            (p + "output.dense.weight", (E, Fd), "w"),
            (p + "output.dense.bias", (E,), "b"),
            (p + "output.LayerNorm.weight", (E,), "ln_w"),
            (p + "output.LayerNorm.bias", (E,), "ln_b"),
        ]
    return out


def make_state_dict(cfg: SynthConfig, seed: int = 0, prefix: str = "", weight_std: float = 0.02,
                    rich: bool = True) -> Dict[str, Tensor]:
    """Seeded random weights.  Matrices follow HF ``_init_weights`` (N(0, 0.02), pad rows zero;
    HF modeling_utils.py:2285-2325).  With ``rich=True`` biases and LayerNorm affine parameters
    are randomised too (HF initialises them to 0 / 1) so that a kernel that drops a bias or a
    gamma cannot pass parity by accident.  numpy's PCG64 stream is platform independent, so the
    GPU box regenerates bit-identical weights without any file travelling."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape, kind in state_dict_keys(cfg, prefix):
        if kind in ("w", "emb", "emb_pad"):
            t = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(weight_std))
            if kind == "emb_pad":
                t[cfg.pad_token_id].zero_()
        elif kind == "b":
            t = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02 if rich else 0.0))
        elif kind == "ln_w":
            t = 1.0 + torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1 if rich else 0.0))
        elif kind == "ln_b":
            t = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1 if rich else 0.0))
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[key] = t
    # buffer the reference registers (ref: recformer/models.py:100)
    sd[prefix + "embeddings.position_ids"] = torch.arange(cfg.max_position_embeddings).expand((1, -1)).clone()
    return sd


def make_batch(cfg: SynthConfig, B: int, L: int, seed: int = 0, ragged: bool = False,
               min_frac: float = 0.5) -> Dict[str, Tensor]:
    """Synthetic batch in the tokenizer's 5-tensor layout (SURVEY.md §8a Spec T / §8d;
    ref: recformer/tokenization.py:64-152).  Items are U[20,96] tokens (3 attrs x <=32), packed
    most-recent-first; ragged rows are right-padded with (ids 1, item_pos max_item-1, type 3,
    mask 0)."""
    rng = np.random.default_rng(seed + 1000)
    ids = np.full((B, L), cfg.pad_token_id, dtype=np.int64)
    tt = np.full((B, L), 3, dtype=np.int64)
    ip = np.full((B, L), cfg.max_item_embeddings - 1, dtype=np.int64)
    am = np.zeros((B, L), dtype=np.int64)
    gm = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        n = L if not ragged else int(rng.integers(max(2, int(L * min_frac)), L + 1))
        if ragged and b == 0:
            n = L  # keep the batch max at L, as `padding` would (tokenization.py:114)
        ids[b, :n] = rng.integers(3, cfg.vocab_size, size=n)
        tt[b, :n] = rng.integers(1, 3, size=n)
        ids[b, 0] = cfg.bos_token_id
        tt[b, 0] = 0
        ip[b, 0] = 0
        pos, item = 1, 1
        while pos < n:
            ln = int(rng.integers(20, 97))
            ip[b, pos:min(n, pos + ln)] = min(item, cfg.max_item_embeddings - 1)
            pos += ln
            item += 1
        am[b, :n] = 1
        gm[b, 0] = 1
    t = lambda a: torch.from_numpy(a)
    return {"input_ids": t(ids), "attention_mask": t(am), "global_attention_mask": t(gm),
            "token_type_ids": t(tt), "item_position_ids": t(ip)}


def make_item_table(N: int, E: int = 768, seed: int = 1) -> Tensor:
    """Independent N(0,1) item table (SURVEY.md §7 hard part 4: a table encoded by a random-init
    encoder is collinear and makes top-k checks vacuous)."""
    rng = np.random.default_rng(seed + 2000)
    return torch.from_numpy(rng.standard_normal((N, E), dtype=np.float32))


# ----------------------------------------------------------------------------------------------
# pretraining weights / batches (ref: recformer/models.py:372-405; LM head = HF:1264-1283 shapes)
# ----------------------------------------------------------------------------------------------
MASK_TOKEN_ID = 50264      # <mask> of the roberta/longformer vocabulary


def lm_head_keys(cfg: SynthConfig) -> List[tuple]:
    E, V = cfg.hidden_size, cfg.vocab_size
    return [("lm_head.bias", (V,), "b"), ("lm_head.dense.weight", (E, E), "w"), ("lm_head.dense.bias", (E,), "b"),
            ("lm_head.layer_norm.weight", (E,), "ln_w"), ("lm_head.layer_norm.bias", (E,), "ln_b"),
            ("lm_head.decoder.weight", (V, E), "w"), ("lm_head.decoder.bias", (V,), "b")]


def make_pretrain_state_dict(cfg: SynthConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Encoder weights under `longformer.` + seeded LM-head weights (same conventions as make_state_dict)."""
    sd = make_state_dict(cfg, seed=seed, prefix="longformer.")
    rng = np.random.default_rng(seed + 4000)
    for key, shape, kind in lm_head_keys(cfg):
        if kind == "w":
            t = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02))
        elif kind == "b":
            t = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02))
        elif kind == "ln_w":
            t = 1.0 + torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1))
        else:
            t = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1))
        sd[key] = t
    return sd


def make_pretrain_batch(cfg: SynthConfig, B: int, La: int, Lb: int, seed: int = 0, mlm_prob: float = 0.15,
                        mask_token_id: Optional[int] = None) -> Dict[str, Tensor]:
    """Synthetic batch of LitWrapper.training_step's layout (ref: recformer/models.py:382-405, collator.py:11-242):
    sequence a (history, ragged, <= La tokens), sequence b (the target item, <= Lb tokens), and their MLM copies
    (15 % of the real non-CLS tokens replaced by <mask>, labels = original id there, -100 elsewhere)."""
    a = make_batch(cfg, B, La, seed=seed, ragged=True)
    b = make_batch(cfg, B, Lb, seed=seed + 7, ragged=True, min_frac=0.25)
    rng = np.random.default_rng(seed + 3000)
    mid = min(MASK_TOKEN_ID, cfg.vocab_size - 1) if mask_token_id is None else mask_token_id
    out = {}
    for tag, d in (("a", a), ("b", b)):
        ids = d["input_ids"].numpy()
        real = d["attention_mask"].numpy().astype(bool)
        real[:, 0] = False
        pick = (rng.random(ids.shape) < mlm_prob) & real
        mlm_ids = np.where(pick, mid, ids)
        labels = np.where(pick, ids, -100)
        for k, v in d.items():
            out[f"{k}_{tag}"] = v
        out[f"mlm_input_ids_{tag}"] = torch.from_numpy(mlm_ids)
        out[f"mlm_labels_{tag}"] = torch.from_numpy(labels)
    return out


# ----------------------------------------------------------------------------------------------
# binary classification head (ref: recformer/models.py:633-660: classifier = 768 -> 384 -> 192 -> 1 Sequential)
# ----------------------------------------------------------------------------------------------
def classifier_keys(cfg: SynthConfig) -> List[tuple]:
    E = cfg.hidden_size
    return [("classifier.0.weight", (E // 2, E), "w"), ("classifier.0.bias", (E // 2,), "b"),
            ("classifier.3.weight", (E // 4, E // 2), "w"), ("classifier.3.bias", (E // 4,), "b"),
            ("classifier.6.weight", (1, E // 4), "w"), ("classifier.6.bias", (1,), "b")]


def make_fraud_state_dict(cfg: SynthConfig, seed: int = 0, head_std: float = 0.05) -> Dict[str, Tensor]:
    """Encoder weights under `longformer.` + a seeded classifier head.  The head's matrices are N(0, 0.05) rather than
    HF's 0.02 so that the logits (three small layers deep) are O(0.1-1) and a parity check on them has teeth."""
    sd = make_state_dict(cfg, seed=seed, prefix="longformer.")
    rng = np.random.default_rng(seed + 5000)
    for key, shape, kind in classifier_keys(cfg):
        std = head_std if kind == "w" else 0.05
        sd[key] = torch.from_numpy(rng.standard_normal(shape, dtype=np.float32) * np.float32(std))
    return sd
