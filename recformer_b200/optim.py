"""Fused AdamW over the engine's flat parameter buffer (ref: optimization.py:7-34 builds torch
AdamW with no weight decay on biases / LayerNorm).  One kernel launch per contiguous parameter
segment (dense encoder weights — which also get their bf16 shadow refreshed in the same pass —,
other matrices, no-decay vectors); frozen parameters are skipped.

Trainable parameters of the wrapped model that live OUTSIDE the encoder's flat buffer (the pretraining `lm_head.*`, a
trainable `item_embedding` made by `init_item_embedding(None)`) are driven by an internal `torch.optim.AdamW` with the
reference's grouping (no decay on biases / LayerNorm weights, ref: optimization.py:12-22), stepped and zeroed together
with the fused segments; `dist.GradSync.finish()` all-reduces their gradients too.  Note that `RecformerForSeqRec`
detaches the item table and the pooled vectors in `similarity_score()` (the table is frozen upstream), so a loss built
from returned SCORES carries no gradient; the trained paths are `forward(labels=...)`."""
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
        self._plan_sig = None
        self._ov = None        # state of an update overlapped with the backward pass (begin_overlap)
        # trainable parameters the flat buffer does not hold
        inside = {id(p) for p in enc.parameters()}
        named_extra = [(k, p) for k, p in model.named_parameters() if id(p) not in inside and p.requires_grad]
        self.extra_params = [p for _, p in named_extra]
        self._extra_opt = None
        if named_extra:
            nd = ("bias", "LayerNorm.weight", "layer_norm.weight")
            groups = [{"params": [p for k, p in named_extra if not any(t in k for t in nd)], "weight_decay": weight_decay},
                      {"params": [p for k, p in named_extra if any(t in k for t in nd)], "weight_decay": 0.0}]
            self._extra_opt = torch.optim.AdamW([g for g in groups if g["params"]], lr=lr, betas=betas, eps=eps)

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

    def _signature(self):
        P = self.engine.params
        return tuple((id(P._named[k]), P._named[k].requires_grad) for k in P._order)

    def zero_grad(self, set_to_none: bool = False):
        P = self.engine.params
        if P.grad is not None:
            P.grad.zero_()
        if self._extra_opt is not None:
            self._extra_opt.zero_grad(set_to_none=True)

    def step_scalars(self, grad_scale: float = 1.0):
        """(lr, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale) of the NEXT step: what step(hp=...) reads from
        device memory when the step is replayed from a CUDA graph (recformer_b200.graph)."""
        t = self.step_count + 1
        b1, b2 = self.betas
        return (float(self.lr), 1.0 - b1 ** t, (1.0 - b2 ** t) ** 0.5, float(grad_scale))

    # -- internals -----------------------------------------------------------------------------
    def _ensure_state(self, frozen_plan: bool) -> None:
        P = self.engine.params
        if P.grad is None:
            raise RuntimeError("FusedAdamW.step() called before any backward pass")
        if self.exp_avg is None or self.exp_avg.data_ptr() == 0 or self.exp_avg.numel() != P.n_total:
            self.exp_avg = torch.zeros_like(P.flat)
            self.exp_avg_sq = torch.zeros_like(P.flat)
            self._plan_sig = None
        if not frozen_plan:              # (a captured step froze its plan; see GraphedTrainStep)
            sig = self._signature()      # requires_grad flips / replaced Parameters since the last step re-plan
            if sig != self._plan_sig:
                self._plan()
                self._plan_sig = sig
        elif self._plan_sig is None:
            self._plan()
            self._plan_sig = self._signature()

    def _update(self, a: int, b: int, decay: bool, shadow: bool, t: int, grad_scale: float, hp, zero: bool,
                wire=None) -> None:
        """One launch over the flat range [a, b)."""
        P = self.engine.params
        b1, b2 = self.betas
        wd = self.weight_decay if decay else 0.0
        sh = P.shadow[a:b] if shadow else None
        if wire is None:
            wire = P.grad_wire      # bf16 all-reduced gradients of this step (dist.GradSync), else None
        if wire is not None:
            ops.adamw_step_bf16grad(P.flat[a:b], wire[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], sh, self.lr, b1, b2,
                                    self.eps, wd, t, grad_scale, hp=hp)
        elif zero:
            ops.adamw_step_zero(P.flat[a:b], P.grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], sh, self.lr, b1, b2,
                                self.eps, wd, t, grad_scale, hp=hp)
        elif hp is not None:
            ops.adamw_step_dev(P.flat[a:b], P.grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], sh, b1, b2, self.eps,
                               wd, hp)
        else:
            ops.adamw_step(P.flat[a:b], P.grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], sh, self.lr, b1, b2,
                           self.eps, wd, t, grad_scale)

    # -- update overlapped with the backward pass ------------------------------------------------
    @torch.no_grad()
    def begin_overlap(self, grad_scale: float = 1.0, hp=None, zero_grads: bool = False, passes: int = 1,
                      sync=None) -> None:
        """Call between the forward and `loss.backward()` of a step whose gradients are complete after this backward
        (no accumulation across calls, no gradient clipping, single GPU): as soon as the engine has enqueued a layer's
        backward, that layer's dense and *_global weights are updated on the engine's aux stream, next to the backward
        kernels of the layers below (AdamW is HBM-bound and small enough to share an SM with a resident GEMM CTA; the
        GEMMs are tensor-bound).  `step()` afterwards updates what is left (embeddings, biases, LayerNorm vectors) and
        joins the aux stream.  The result is bit-identical to a plain `step()`: same kernels on the same inputs.
        zero_grads: every update also clears the gradient range it consumed (no zero_grad() memset next step).
        passes: encoder backward passes per step (a layer is final after the last one).
        sync: the step's dist.GradSync (data-parallel).  The updates then follow the all-reduce of each layer bucket
        instead of each layer's backward: the aux stream waits for the bucket's collectives and updates the bucket's
        ranges from the reduced bf16 wire gradients while the backward of the layers below is still running; pass
        grad_scale = 1 / world."""
        P = self.engine.params
        if P.grad_wire is not None:
            raise RuntimeError("FusedAdamW.begin_overlap: a previous GradSync.finish() was never consumed by step()")
        P.prepare_grads()
        self._ensure_state(frozen_plan=hp is not None)
        self._ov = {"gs": float(grad_scale), "hp": hp, "zero": bool(zero_grads), "covered": [], "seen": {},
                    "passes": int(passes), "prev": self.engine.grad_hook, "t": self.step_count + 1, "last": None,
                    "sync": sync}
        if sync is None:
            self.engine.grad_hook = self._on_layer
        else:
            if zero_grads:
                raise RuntimeError("FusedAdamW.begin_overlap: zero_grads is for fp32 gradients, not the all-reduced wire")
            sync.bucket_hook = self._on_bucket

    def _on_bucket(self, ranges, works) -> None:
        """dist.GradSync launched the all-reduces of a layer bucket (flat `ranges`)."""
        ov = self._ov
        if not ov["sync"]._active():
            return
        eng = self.engine
        device = eng.params.flat.device
        wire = ov["sync"]._wire(eng.params.grad)
        aux = eng.aux_stream(device)
        with torch.cuda.stream(aux):
            for w in works:
                w.wait()                       # the aux stream waits for the collectives (not the host)
            for r0, r1 in ranges:
                for (a, b, decay, shadow) in self._segments:
                    lo, hi = max(a, r0), min(b, r1)
                    if lo < hi:
                        self._update(lo, hi, decay, shadow, ov["t"], ov["gs"], ov["hp"], False,
                                     wire=wire if wire is not None else None)
                        ov["covered"].append((lo, hi))
            done = eng.event(device, "opt.bucket.done", len(ov["covered"]))
            done.record(aux)
        ov["last"] = done

    def _on_layer(self, layer: int) -> None:
        ov = self._ov
        if ov["prev"] is not None:
            ov["prev"](layer)
        ov["seen"][layer] = ov["seen"].get(layer, 0) + 1
        if ov["seen"][layer] < ov["passes"]:
            return
        eng = self.engine
        device = eng.params.flat.device

        def run():
            for r0, r1 in eng.params.layer_ranges(layer):
                for (a, b, decay, shadow) in self._segments:
                    lo, hi = max(a, r0), min(b, r1)
                    if lo < hi:
                        self._update(lo, hi, decay, shadow, ov["t"], ov["gs"], ov["hp"], ov["zero"])
                        ov["covered"].append((lo, hi))

        ov["last"] = eng.fork_aux(device, "opt", layer, run)

    @staticmethod
    def _subtract(a: int, b: int, covered):
        """[a, b) minus the (disjoint) covered ranges."""
        out, pos = [], a
        for lo, hi in sorted(covered):
            if hi <= pos or lo >= b:
                continue
            if lo > pos:
                out.append((pos, lo))
            pos = max(pos, hi)
        if pos < b:
            out.append((pos, b))
        return out

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, wait_other=None, hp=None, zero_grads: bool = False):
        """wait_other: optional callable (GradSync.finish(defer_tail=True)) that is invoked right before the
        segment holding the embedding tables is updated — the dense and no-decay segments run while the
        embedding gradients are still being all-reduced.
        hp: optional fp32 device tensor holding step_scalars(); the kernels then take lr, the bias corrections
        and grad_scale from it instead of from their arguments (graph capture / replay).
        zero_grads: the update also clears the fp32 gradients it consumed (see begin_overlap)."""
        P = self.engine.params
        ov, self._ov = self._ov, None
        if ov is not None:
            if ov["sync"] is None:
                self.engine.grad_hook = ov["prev"]
            else:
                ov["sync"].bucket_hook = None
            grad_scale, hp, zero_grads = ov["gs"], ov["hp"], ov["zero"]
        self._ensure_state(frozen_plan=hp is not None)
        if self._extra_opt is not None:
            if hp is not None:
                raise RuntimeError("FusedAdamW: parameters outside the encoder's flat buffer cannot be stepped from a "
                                   "captured graph (hp=...); step them eagerly")
            for g in self._extra_opt.param_groups:
                g["lr"] = self.lr
            if grad_scale != 1.0:
                for p in self.extra_params:
                    if p.grad is not None:
                        p.grad.mul_(grad_scale)
            self._extra_opt.step()
        self.step_count += 1
        segs = list(self._segments)
        if wait_other is not None:      # embeddings (+ *_global weights) = the decay segment without a bf16 shadow: last
            segs.sort(key=lambda sgm: (sgm[2] and not sgm[3]))
        covered = ov["covered"] if ov is not None else []
        for (a, b, decay, shadow) in segs:
            if wait_other is not None and decay and not shadow:
                wait_other()
                wait_other = None
            for lo, hi in (self._subtract(a, b, covered) if covered else [(a, b)]):
                self._update(lo, hi, decay, shadow, self.step_count, grad_scale, hp, zero_grads and P.grad_wire is None)
        if wait_other is not None:
            wait_other()
        if ov is not None and ov["last"] is not None:      # join the aux stream: every parameter is updated on return
            torch.cuda.current_stream(P.flat.device).wait_event(ov["last"])
        P.grad_wire = None            # consumed: a later step without a GradSync must not read stale gradients
        P.mark_shadow_fresh(by_optimizer=True)

    def untouched_ranges(self):
        """Flat ranges no update ever writes (frozen parameters): what a step that relies on zero_grads=True still has
        to clear itself."""
        self._ensure_state(frozen_plan=False)
        return self._subtract(0, self.engine.params.n_total, [(a, b) for (a, b, _, _) in self._segments])
