"""Ranking metrics (ref: utils.py:76-107; SURVEY.md §8a Spec R).

`Ranker` keeps the reference semantics on a dense (B,N) score tensor.  `TopKRanker` computes the
@k metrics from the fused scorer's output (top-k scores + label score) without ever holding the
(B,N) logits: with c = #{t_i > s*}, rank < k <=> c < k and then rank = c."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

MAX_VAL = 1e4


class Ranker(nn.Module):
    def __init__(self, metrics_ks: Sequence[int]):
        super().__init__()
        self.ks = list(metrics_ks)
        self.ce = nn.CrossEntropyLoss()

    def forward(self, scores: torch.Tensor, labels: torch.Tensor) -> List[float]:
        # ref utils.py:83 squeezes; a final batch of ONE user then has a 0-d target and the reference's CE raises
        # (caught upstream as loss 0.0, utils.py:84-89).  view(-1) keeps the (B,) target for every B.
        labels = labels.view(-1).long()
        loss = self.ce(scores, labels).item()
        predicts = scores[torch.arange(scores.size(0), device=scores.device), labels].unsqueeze(-1)
        valid_length = (scores > -MAX_VAL).sum(-1).float()
        rank = (predicts < scores).sum(-1).float()
        res = []
        for k in self.ks:
            indicator = (rank < k).float()
            res.append(((1 / torch.log2(rank + 2)) * indicator).mean().item())
            res.append(indicator.mean().item())
        res.append((1 / (rank + 1)).mean().item())
        res.append((1 - (rank / valid_length)).mean().item())
        return res + [loss]


class TopKRanker:
    """NDCG@k / Recall@k for k <= K from (B,K) descending top-K scores and (B,) label scores."""

    def __init__(self, metrics_ks: Sequence[int]):
        self.ks = list(metrics_ks)

    def __call__(self, topk_scores: torch.Tensor, label_scores: torch.Tensor) -> List[float]:
        res = []
        for k in self.ks:
            if k > topk_scores.shape[1]:
                raise ValueError(f"k={k} exceeds the {topk_scores.shape[1]} scores kept")
            c = (topk_scores[:, :k] > label_scores[:, None]).sum(-1).float()
            ind = (c < k).float()
            res.append(((1 / torch.log2(c + 2)) * ind).mean().item())
            res.append(ind.mean().item())
        return res
