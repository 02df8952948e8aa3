"""CPU oracle for the Recformer encoder + scoring hot path.

TEST INFRASTRUCTURE ONLY.  This module is a plain-PyTorch (CPU, fp32) restatement of the
reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
``recformer_b200`` never does (it fails loudly when its CUDA library is missing).

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the *unmodified* reference ``recformer/models.py`` run in the build
container under the import shim ``oracle/ref_shim.py``; the resulting fixtures live in
``tests/golden/`` together with the script that made them (``tests/golden/make_goldens.py``)
and ``tests/test_oracle_golden.py`` re-checks the oracle against them on every run.

Every function cites the reference lines it restates.  ``ref:`` is /root/reference,
``HF:`` is transformers/models/longformer/modeling_longformer.py (5.5.0 in this image; the
reference pins 4.28.0 whose Longformer arithmetic is the same).

The model is expressed functionally over a ``state_dict`` (same keys as the reference,
SURVEY.md §8b) so autograd through it yields oracle gradients for the backward kernels.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# config + deterministic synthetic weights / inputs: kept on neutral ground (tools/synthetic.py) so that bench.py's
# product arm and the tools can build their inputs without importing this checker; re-exported under the old names.
# ----------------------------------------------------------------------------------------------
from tools.synthetic import (SynthConfig as OracleConfig, MASK_TOKEN_ID, state_dict_keys, make_state_dict,  # noqa: E402,F401
                             make_batch, make_item_table, lm_head_keys, make_pretrain_state_dict, make_pretrain_batch,
                             classifier_keys, make_fraud_state_dict)


# ----------------------------------------------------------------------------------------------
# host-side mask / id preparation
# ----------------------------------------------------------------------------------------------
def create_position_ids_from_input_ids(input_ids: Tensor, padding_idx: int) -> Tensor:
    """ref: recformer/models.py:68-79 — cumsum(ids != pad) * (ids != pad) + pad."""
    mask = input_ids.ne(padding_idx).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx


def merge_to_attention_mask(attention_mask: Optional[Tensor], global_attention_mask: Tensor) -> Tensor:
    """ref: recformer/models.py:262-272 — 0 pad / 1 local / 2 global."""
    if attention_mask is not None:
        return attention_mask * (global_attention_mask + 1)
    return global_attention_mask + 1


def pad_to_window_size(cfg: OracleConfig, input_ids, attention_mask, token_type_ids, position_ids,
                       item_position_ids):
    """ref: recformer/models.py:210-260 — right-pad to a multiple of max(attention_window):
    ids<-pad, position_ids<-pad, item_position_ids<-pad (sic, :244), mask<-0, token_type<-0."""
    w = max(cfg.attention_window)
    assert w % 2 == 0
    L = input_ids.shape[1]
    padding_len = (w - L % w) % w
    if padding_len > 0:
        input_ids = F.pad(input_ids, (0, padding_len), value=cfg.pad_token_id)
        if position_ids is not None:
            position_ids = F.pad(position_ids, (0, padding_len), value=cfg.pad_token_id)
        if item_position_ids is not None:
            item_position_ids = F.pad(item_position_ids, (0, padding_len), value=cfg.pad_token_id)
        attention_mask = F.pad(attention_mask, (0, padding_len), value=0)
        token_type_ids = F.pad(token_type_ids, (0, padding_len), value=0)
    return padding_len, input_ids, attention_mask, token_type_ids, position_ids, item_position_ids


# ----------------------------------------------------------------------------------------------
# embeddings (ref: recformer/models.py:108-138)
# ----------------------------------------------------------------------------------------------
def embeddings_forward(sd: Dict[str, Tensor], cfg: OracleConfig, input_ids, token_type_ids,
                       item_position_ids, position_ids=None, prefix: str = "", dropout_p: float = 0.0):
    if position_ids is None:
        position_ids = create_position_ids_from_input_ids(input_ids, cfg.pad_token_id)
    p = prefix + "embeddings."
    x = (F.embedding(input_ids, sd[p + "word_embeddings.weight"])
         + F.embedding(position_ids, sd[p + "position_embeddings.weight"])
         + F.embedding(token_type_ids, sd[p + "token_type_embeddings.weight"])
         + F.embedding(item_position_ids, sd[p + "item_position_embeddings.weight"]))
    x = F.layer_norm(x, (cfg.hidden_size,), sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"],
                     cfg.layer_norm_eps)
    return F.dropout(x, dropout_p, training=dropout_p > 0)


# ----------------------------------------------------------------------------------------------
# Longformer self-attention, dense-mask restatement (SURVEY.md §8a Spec A; HF:481-639,963-1056)
# ----------------------------------------------------------------------------------------------
def self_attention_forward(sd: Dict[str, Tensor], cfg: OracleConfig, layer: int, hidden: Tensor,
                           mask012: Tensor, prefix: str = "") -> Tensor:
    """hidden (B,L,E); mask012 (B,L) with 0 = padding, 1 = local, 2 = global.

    Non-global query i attends keys {j: |i-j| <= w, j valid, j not global} U {j global}
    (HF:523 removes global keys from the band, HF:558-568 re-adds them as prepended columns);
    padded query rows output exactly zero (HF:578).  A global query g uses the *_global
    projections over every valid key (HF:963-1056) and overwrites row g (HF:626).  Softmax in
    fp32.  Executed here with dense (L,L) masks — O(L^2) but obviously equal to the band."""
    B, L, E = hidden.shape
    H = cfg.num_attention_heads
    D = E // H
    w = cfg.attention_window[layer] // 2
    p = f"{prefix}encoder.layer.{layer}.attention.self."
    lin = lambda n, x: F.linear(x, sd[p + n + ".weight"], sd[p + n + ".bias"])

    valid = mask012 > 0            # (B,L)
    glob = mask012 > 1
    q = lin("query", hidden) / math.sqrt(D)      # HF:503,513
    k = lin("key", hidden)
    v = lin("value", hidden)
    heads = lambda t: t.view(B, L, H, D).transpose(1, 2)   # (B,H,L,D)
    qh, kh, vh = heads(q), heads(k), heads(v)

    idx = torch.arange(L)
    band = (idx[:, None] - idx[None, :]).abs() <= w                     # (L,L)
    allowed = (band[None] & (valid & ~glob)[:, None, :]) | glob[:, None, :]     # (B,L,L)
    scores = torch.matmul(qh, kh.transpose(-1, -2))                     # (B,H,L,L)
    scores = scores.masked_fill(~allowed[:, None], float("-inf"))
    probs = torch.softmax(scores.float(), dim=-1)
    probs = torch.nan_to_num(probs, nan=0.0)
    probs = probs.masked_fill(~valid[:, None, :, None], 0.0)            # HF:578
    out = torch.matmul(probs, vh)                                       # (B,H,L,D)

    if bool(glob.any()):
        qg = heads(lin("query_global", hidden) / math.sqrt(D))          # HF:979,986
        kg = heads(lin("key_global", hidden))
        vg = heads(lin("value_global", hidden))
        sg = torch.matmul(qg, kg.transpose(-1, -2))
        sg = sg.masked_fill(~valid[:, None, None, :], float("-inf"))    # HF:1023-1026
        pg = torch.softmax(sg.float(), dim=-1)
        og = torch.matmul(pg, vg)
        out = torch.where(glob[:, None, :, None], og, out)              # HF:615-626
    return out.transpose(1, 2).reshape(B, L, E)


def gelu_erf(x: Tensor) -> Tensor:
    """HF ACT2FN['gelu'] — exact erf form (HF:1103-1115)."""
    return F.gelu(x)


def layer_forward(sd, cfg: OracleConfig, layer: int, hidden: Tensor, mask012: Tensor, prefix: str = ""):
    """HF:1133-1171 — post-LN blocks: h1 = LN(dense(attn)+h); out = LN(W2 gelu(W1 h1) + h1)."""
    p = f"{prefix}encoder.layer.{layer}."
    E = cfg.hidden_size
    a = self_attention_forward(sd, cfg, layer, hidden, mask012, prefix)
    a = F.linear(a, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
    h1 = F.layer_norm(a + hidden, (E,), sd[p + "attention.output.LayerNorm.weight"],
                      sd[p + "attention.output.LayerNorm.bias"], cfg.layer_norm_eps)        # HF:1067-1071
    u = F.linear(h1, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"])
    g = gelu_erf(u)                                                                         # HF:1112-1115
    d = F.linear(g, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return F.layer_norm(d + h1, (E,), sd[p + "output.LayerNorm.weight"],
                        sd[p + "output.LayerNorm.bias"], cfg.layer_norm_eps)                # HF:1126-1130


# ----------------------------------------------------------------------------------------------
# RecformerModel.forward (ref: recformer/models.py:274-356)
# ----------------------------------------------------------------------------------------------
def model_forward(sd: Dict[str, Tensor], cfg: OracleConfig, input_ids: Tensor,
                  attention_mask: Optional[Tensor] = None, global_attention_mask: Optional[Tensor] = None,
                  token_type_ids: Optional[Tensor] = None, item_position_ids: Optional[Tensor] = None,
                  position_ids: Optional[Tensor] = None, prefix: str = "", num_layers: Optional[int] = None):
    """Returns (last_hidden_state (B,L,E), pooler_output (B,E))."""
    B, L = input_ids.shape
    if attention_mask is None:
        attention_mask = torch.ones(B, L, dtype=torch.long)
    if token_type_ids is None:
        token_type_ids = torch.zeros(B, L, dtype=torch.long)
    if global_attention_mask is not None:
        attention_mask = merge_to_attention_mask(attention_mask, global_attention_mask)
    padding_len, input_ids, attention_mask, token_type_ids, position_ids, item_position_ids = \
        pad_to_window_size(cfg, input_ids, attention_mask, token_type_ids, position_ids, item_position_ids)
    # the additive mask of models.py:327-329 only carries sign information into HF:1188-1190
    mask012 = attention_mask
    h = embeddings_forward(sd, cfg, input_ids, token_type_ids, item_position_ids, position_ids, prefix)
    nl = cfg.num_hidden_layers if num_layers is None else num_layers
    for i in range(nl):
        h = layer_forward(sd, cfg, i, h, mask012, prefix)
    if padding_len > 0:
        h = h[:, : h.shape[1] - padding_len]          # HF:1228
    if cfg.pooler_type == "cls":
        pooled = h[:, 0]                              # ref: recformer/models.py:165
    elif cfg.pooler_type == "avg":
        am = attention_mask[:, : h.shape[1]].to(h.dtype)
        pooled = (h * am.unsqueeze(-1)).sum(1) / am.sum(-1).unsqueeze(-1)
    else:
        raise NotImplementedError
    return h, pooled


# ----------------------------------------------------------------------------------------------
# scoring (SURVEY.md §8a Spec S; ref: recformer/models.py:358-369,539-545)
# ----------------------------------------------------------------------------------------------
def similarity(x: Tensor, y: Tensor, temp: float) -> Tensor:
    """nn.CosineSimilarity(dim=-1)(x, y) / temp on broadcastable x (B,1,E), y (1|B,N,E).
    ATen: x/max(|x|,1e-8) . y/max(|y|,1e-8).  Written as normalise -> matmul, which the survey
    probed equal to the reference's broadcast form to 7e-7."""
    xn = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    yn = y / y.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    if y.shape[0] == 1 and x.shape[1] == 1:      # whole-table form: (B,E) @ (E,N)
        return torch.matmul(xn[:, 0], yn[0].T) / temp
    return (xn * yn).sum(-1) / temp              # per-row candidates (B,C,E)


def similarity_score(pooled: Tensor, item_embedding: Tensor, temp: float, candidates: Optional[Tensor] = None):
    """ref: recformer/models.py:539-545."""
    if candidates is None:
        cand = item_embedding.unsqueeze(0)
    else:
        cand = F.embedding(candidates, item_embedding)
    return similarity(pooled.unsqueeze(1), cand, temp)


def seqrec_forward(sd, cfg: OracleConfig, batch: Dict[str, Tensor], item_embedding: Tensor,
                   labels: Optional[Tensor] = None, candidates: Optional[Tensor] = None,
                   prefix: str = "longformer."):
    """ref: recformer/models.py:547-599 — scores (B,N) without labels; CE loss with labels
    (full softmax when finetune_negative_sample_size <= 0; sampled negatives are drawn by the
    caller and passed as `candidates` with the label in column 0, :593-597)."""
    _, pooled = model_forward(sd, cfg, prefix=prefix, **batch)
    if labels is None:
        return similarity_score(pooled, item_embedding, cfg.temp, candidates)
    if candidates is None:
        logits = similarity_score(pooled, item_embedding, cfg.temp)
        return F.cross_entropy(logits, labels)
    logits = similarity_score(pooled, item_embedding, cfg.temp, candidates)
    return F.cross_entropy(logits, torch.zeros_like(labels))


# ----------------------------------------------------------------------------------------------
# binary classification head (ref: recformer/models.py:601-713; caller finetune_classification.py)
# ----------------------------------------------------------------------------------------------
def focal_loss(inputs: Tensor, targets: Tensor, alpha: Optional[float] = 1, gamma: float = 2,
               pos_weight: Optional[Tensor] = None) -> Tensor:
    """ref: recformer/models.py:611-631 — p = sigmoid(x); p_t = p t + (1-p)(1-t); loss = alpha_t (1-p_t)^gamma BCE(x, t)
    with alpha_t = alpha t + (1-alpha)(1-t) (skipped when alpha is None); mean over the batch."""
    p = torch.sigmoid(inputs)
    # BCE with logits and pos_weight w: -(w t log p + (1-t) log(1-p)), written with logsigmoid for stability
    w = 1.0 if pos_weight is None else pos_weight
    ce = -(w * targets * F.logsigmoid(inputs) + (1 - targets) * F.logsigmoid(-inputs))
    p_t = p * targets + (1 - p) * (1 - targets)
    out = (1 - p_t) ** gamma * ce
    if alpha is not None:
        out = (alpha * targets + (1 - alpha) * (1 - targets)) * out
    return out.mean()


def fraud_head(sd: Dict[str, Tensor], pooled: Tensor) -> Tensor:
    """ref: recformer/models.py:643-651,697-699 in eval mode (the three dropouts are identities): Linear/ReLU ->
    Linear/ReLU -> Linear -> squeeze(-1).  Returns logits (B,)."""
    h = torch.relu(F.linear(pooled, sd["classifier.0.weight"], sd["classifier.0.bias"]))
    h = torch.relu(F.linear(h, sd["classifier.3.weight"], sd["classifier.3.bias"]))
    return F.linear(h, sd["classifier.6.weight"], sd["classifier.6.bias"]).squeeze(-1)


def bce_with_logits(logits: Tensor, labels: Tensor, pos_weight: float = 1.0) -> Tensor:
    """ref: recformer/models.py:702-708 — nn.BCEWithLogitsLoss(pos_weight=w)(logits, labels.float()):
    mean of -(w t log sigmoid(x) + (1 - t) log sigmoid(-x))."""
    t = labels.float()
    return -(pos_weight * t * F.logsigmoid(logits) + (1 - t) * F.logsigmoid(-logits)).mean()


def fraud_forward(sd: Dict[str, Tensor], cfg: OracleConfig, batch: Dict[str, Tensor], labels: Optional[Tensor] = None,
                  pos_weight: float = 1.0):
    """ref: recformer/models.py:677-713 — encoder -> pooled -> head; returns (loss or None, logits (B,))."""
    _, pooled = model_forward(sd, cfg, prefix="longformer.", **batch)
    logits = fraud_head(sd, pooled)
    return (None if labels is None else bce_with_logits(logits, labels, pos_weight)), logits


# ----------------------------------------------------------------------------------------------
# pretraining step (ref: recformer/models.py:372-520; LM head = HF:1264-1283 LongformerLMHead)
# ----------------------------------------------------------------------------------------------
def lm_head_forward(sd: Dict[str, Tensor], cfg: "OracleConfig", features: Tensor) -> Tensor:
    """HF:1275-1283: decoder(layer_norm(gelu_erf(dense(x)))); the decoder has its own bias (`lm_head.bias` is
    an unused parameter in transformers 5.5.0's copy)."""
    x = F.linear(features, sd["lm_head.dense.weight"], sd["lm_head.dense.bias"])
    x = gelu_erf(x)
    x = F.layer_norm(x, (cfg.hidden_size,), sd["lm_head.layer_norm.weight"], sd["lm_head.layer_norm.bias"],
                     cfg.layer_norm_eps)
    return F.linear(x, sd["lm_head.decoder.weight"], sd["lm_head.decoder.bias"])


def pretrain_forward(sd: Dict[str, Tensor], cfg: "OracleConfig", batch: Dict[str, Tensor], mlm_weight: float = 0.1,
                     gathered_z: Optional[tuple] = None):
    """ref: recformer/models.py:382-520 (single process; `gathered_z` = (z1_all, z2_all, rank) emulates the
    dist.all_gather branch :475-490 with this rank's slot replaced by the live tensors).
    Returns (loss, cos_sim, correct_num)."""
    enc = lambda tag, ids: model_forward(
        sd, cfg, ids, attention_mask=batch[f"attention_mask_{tag}"],
        global_attention_mask=batch[f"global_attention_mask_{tag}"], token_type_ids=batch[f"token_type_ids_{tag}"],
        item_position_ids=batch[f"item_position_ids_{tag}"], prefix="longformer.")
    _, z1 = enc("a", batch["input_ids_a"])
    _, z2 = enc("b", batch["input_ids_b"])
    if gathered_z is not None:
        z1_all, z2_all, rank = gathered_z
        B = z1.shape[0]
        z1 = torch.cat([z1_all[: rank * B], z1, z1_all[(rank + 1) * B:]], 0)
        z2 = torch.cat([z2_all[: rank * B], z2, z2_all[(rank + 1) * B:]], 0)
    cos_sim = similarity(z1.unsqueeze(1), z2.unsqueeze(0), cfg.temp)
    labels = torch.arange(cos_sim.shape[0])
    loss = F.cross_entropy(cos_sim, labels)
    correct = (cos_sim.argmax(1) == labels).sum()
    for tag in ("a", "b"):
        if batch.get(f"mlm_input_ids_{tag}") is not None:
            hidden, _ = enc(tag, batch[f"mlm_input_ids_{tag}"])
            scores = lm_head_forward(sd, cfg, hidden)
            loss = loss + mlm_weight * F.cross_entropy(scores.view(-1, cfg.vocab_size),
                                                       batch[f"mlm_labels_{tag}"].reshape(-1))
    return loss, cos_sim, correct


# ----------------------------------------------------------------------------------------------
# metrics (SURVEY.md §8a Spec R; ref: utils.py:76-107)
# ----------------------------------------------------------------------------------------------
MAX_VAL = 1e4


def ranker(scores: Tensor, labels: Tensor, ks: Sequence[int] = (10, 50)) -> List[float]:
    """[NDCG@k, Recall@k for k in ks] + [MRR, AUC, CE] — rank counts strictly greater scores."""
    labels = labels.reshape(-1)
    loss = F.cross_entropy(scores, labels).item()
    predicts = scores[torch.arange(scores.size(0)), labels].unsqueeze(-1)
    valid_length = (scores > -MAX_VAL).sum(-1).float()
    rank = (predicts < scores).sum(-1).float()
    res = []
    for k in ks:
        ind = (rank < k).float()
        res.append(((1 / torch.log2(rank + 2)) * ind).mean().item())
        res.append(ind.mean().item())
    res.append((1 / (rank + 1)).mean().item())
    res.append((1 - (rank / valid_length)).mean().item())
    return res + [loss]


def topk_metrics(topk_scores: Tensor, label_scores: Tensor, k: int = 10):
    """Recall@k / NDCG@k from (B,k) descending top-k scores and (B,) label scores: with
    c = #{t_i > s*}, rank < k <=> c < k and then rank = c (SURVEY.md §8a Spec R)."""
    c = (topk_scores[:, :k] > label_scores[:, None]).sum(-1).float()
    ind = (c < k).float()
    return ((1 / torch.log2(c + 2)) * ind).mean().item(), ind.mean().item()


# ----------------------------------------------------------------------------------------------
# tokenizer batch layout (SURVEY.md §8a Spec T; ref: recformer/tokenization.py:64-159)
# ----------------------------------------------------------------------------------------------
def tokenizer_encode(cfg: OracleConfig, items: list) -> Dict[str, list]:
    """`encode(items, encode_item=False)`: items are (input_ids, token_type_ids) pairs, oldest
    first; output is <s> + most-recent-first, truncated to max_token_num."""
    items = items[::-1][: cfg.max_item_embeddings - 1]
    input_ids, item_position_ids, token_type_ids = [cfg.bos_token_id], [0], [0]
    for item_idx, (ids, tts) in enumerate(items):
        input_ids += list(ids)
        token_type_ids += list(tts)
        item_position_ids += [item_idx + 1] * len(ids)
    n = cfg.max_token_num
    input_ids, item_position_ids, token_type_ids = input_ids[:n], item_position_ids[:n], token_type_ids[:n]
    gm = [0] * len(input_ids)
    gm[0] = 1
    return {"input_ids": input_ids, "item_position_ids": item_position_ids, "token_type_ids": token_type_ids,
            "attention_mask": [1] * len(input_ids), "global_attention_mask": gm}


def tokenizer_padding(cfg: OracleConfig, item_batch: List[Dict[str, list]], pad_to_max: bool = False):
    max_length = cfg.max_token_num if pad_to_max else max(len(x["input_ids"]) for x in item_batch)
    out = {k: [] for k in ("input_ids", "item_position_ids", "token_type_ids", "attention_mask",
                           "global_attention_mask")}
    fill = {"input_ids": cfg.pad_token_id, "item_position_ids": cfg.max_item_embeddings - 1,
            "token_type_ids": 3, "attention_mask": 0, "global_attention_mask": 0}
    for x in item_batch:
        n = max_length - len(x["input_ids"])
        for k in out:
            out[k].append(list(x[k]) + [fill[k]] * n)
    return out


def tokenizer_batch_encode(cfg: OracleConfig, item_batch: list, pad_to_max: bool = False):
    return tokenizer_padding(cfg, [tokenizer_encode(cfg, items) for items in item_batch], pad_to_max)
