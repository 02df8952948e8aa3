"""Fused AdamW over the engine's flat parameter buffer (ref: optimization.py:7-34 builds torch
AdamW with no weight decay on biases / LayerNorm).  One kernel launch per contiguous parameter
segment (dense encoder weights — which also get their bf16 shadow refreshed in the same pass —,
other matrices, no-decay vectors); frozen parameters are skipped."""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops


class FusedAdamW:
    def __init__(self, model, lr: float = 5e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        enc = getattr(model, "longformer", model)
        self.engine = enc._engine
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self._segments: List[Tuple[int, int, bool, bool]] = []

    def _plan(self):
        P = self.engine.params
        segs = []
        cur = None
        for k in P._order:
            p = P._named[k]
            o, n = P.offsets[k], p.numel()
            decay = o < P.n_decay
            shadow = o < P.n_dense
            if not p.requires_grad:
                cur = None
                continue
            if cur is not None and cur[1] == o and cur[2] == decay and cur[3] == shadow:
                cur[1] = o + n
            else:
                cur = [o, o + n, decay, shadow]
                segs.append(cur)
        self._segments = [tuple(s) for s in segs]

    def zero_grad(self, set_to_none: bool = False):
        P = self.engine.params
        if P.grad is not None:
            P.grad.zero_()

    def step_scalars(self, grad_scale: float = 1.0):
        """(lr, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale) of the NEXT step: what step(hp=...) reads from
        device memory when the step is replayed from a CUDA graph (recformer_b200.graph)."""
        t = self.step_count + 1
        b1, b2 = self.betas
        return (float(self.lr), 1.0 - b1 ** t, (1.0 - b2 ** t) ** 0.5, float(grad_scale))

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, wait_other=None, hp=None):
        """wait_other: optional callable (GradSync.finish(defer_tail=True)) that is invoked right before the
        segment holding the embedding tables is updated — the dense and no-decay segments run while the
        embedding gradients are still being all-reduced.
        hp: optional fp32 device tensor holding step_scalars(); the kernels then take lr, the bias corrections
        and grad_scale from it instead of from their arguments (graph capture / replay)."""
        P = self.engine.params
        if P.grad is None:
            raise RuntimeError("FusedAdamW.step() called before any backward pass")
        if self.exp_avg is None or self.exp_avg.data_ptr() == 0 or self.exp_avg.numel() != P.n_total:
            self.exp_avg = torch.zeros_like(P.flat)
            self.exp_avg_sq = torch.zeros_like(P.flat)
            self._plan()
        self.step_count += 1
        b1, b2 = self.betas
        segs = list(self._segments)
        if wait_other is not None:      # embeddings (+ *_global weights) = the decay segment without a bf16 shadow: last
            segs.sort(key=lambda sgm: (sgm[2] and not sgm[3]))
        for (a, b, decay, shadow) in segs:
            if wait_other is not None and decay and not shadow:
                wait_other()
                wait_other = None
            if hp is not None:
                ops.adamw_step_dev(P.flat[a:b], P.grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b],
                                   P.shadow[a:b] if shadow else None, b1, b2, self.eps,
                                   self.weight_decay if decay else 0.0, hp)
                continue
            ops.adamw_step(P.flat[a:b], P.grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b],
                           P.shadow[a:b] if shadow else None, self.lr, b1, b2, self.eps,
                           self.weight_decay if decay else 0.0, self.step_count, grad_scale)
        if wait_other is not None:
            wait_other()
        P.mark_shadow_fresh(by_optimizer=True)
