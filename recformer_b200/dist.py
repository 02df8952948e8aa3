"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL; gloo for CPU tests).

The hot path shards in two ways:
  * full-catalogue scoring: the item table is sharded by contiguous item-id ranges; each rank
    runs the fused cosine-GEMM + top-k on its shard and the per-rank (k scores, k ids, label
    score) are exchanged with ONE all-gather, then merged locally (rf_topk_merge) — or, with a
    PeerTopkExchange, stored by the scorer straight into every rank's symmetric buffer over NVLink peer
    memory (no NCCL launch: scorer -> barrier -> merge);
  * data-parallel training: replicas; the flat gradient buffer is all-reduced in one call
    (allreduce_gradients) or in layer buckets overlapped with the backward pass (GradSync, bf16 wire format).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous id range [lo, hi) owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_topk(scores: torch.Tensor, ids: torch.Tensor, label_score: torch.Tensor) -> torch.Tensor:
    """(B,k) fp32, (B,k) int32, (B,) fp32 -> one (B, 2k+1) fp32 buffer (ids bit-cast) so that the
    exchange is a single collective."""
    return torch.cat([scores, ids.view(torch.float32), label_score[:, None]], dim=1).contiguous()


def unpack_topk(buf: torch.Tensor, k: int):
    """(W, B, 2k+1) -> scores (W,B,k) fp32, ids (W,B,k) int32, label scores (W,B) fp32."""
    return (buf[..., :k].contiguous(), buf[..., k:2 * k].contiguous().view(torch.int32), buf[..., 2 * k].contiguous())


def merge_topk_reference(scores: torch.Tensor, ids: torch.Tensor, label_scores: torch.Tensor, k: int):
    """Host/torch restatement of rf_topk_merge used by the CPU (gloo) tests: global top-k over W
    lists of k with ties broken towards the lower id; label score = max over ranks."""
    W, B, _ = scores.shape
    s = scores.permute(1, 0, 2).reshape(B, W * k)
    i = ids.permute(1, 0, 2).reshape(B, W * k).to(torch.int64)
    order = torch.argsort(i, dim=1, stable=True)                 # ids ascending first ...
    s, i = torch.gather(s, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)  # ... then stable by score
    s, i = torch.gather(s, 1, order)[:, :k], torch.gather(i, 1, order)[:, :k]
    return s, i.to(torch.int32), label_scores.max(dim=0).values


def all_gather_topk(scores, ids, label_score, group=None):
    """All-gather the packed per-rank top-k; returns (W,B,k) scores, (W,B,k) ids, (W,B) labels."""
    k = scores.shape[1]
    packed = pack_topk(scores, ids, label_score)
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out.view(-1), packed.view(-1), group=group)
    return unpack_topk(out, k)


def sharded_topk(model, pooled: torch.Tensor, k: int = 10, labels: Optional[torch.Tensor] = None,
                 id_base: int = 0, group=None, exchange: Optional["PeerTopkExchange"] = None):
    """Global top-k when `model.item_embedding` holds only this rank's shard (ids offset by
    id_base).  Three launches + one collective: the fused scorer writes this rank's packed (B, 2k+1) result
    (rf_cosine_topk_packed), ONE all-gather exchanges it, rf_topk_merge_packed reads the gathered buffer in place;
    the result is identical on every rank and bit-identical to the unsharded run."""
    from . import ops
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return model.topk(pooled, k=k, labels=labels, id_base=id_base)
    if exchange is not None:      # peer-memory exchange fused into the scorer (see PeerTopkExchange)
        return exchange.topk(model, pooled, labels=labels, id_base=id_base)
    xn = ops.normalize_rows(pooled.contiguous())
    packed = ops.cosine_topk_packed(xn, model.normalized_items(), model.config.temp, k=k, id_base=id_base, labels=labels)
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out.view(-1), packed.view(-1), group=group)
    return ops.topk_merge_packed(out, k)


class PeerTopkExchange:
    """The all-gather of sharded scoring fused into the scorer: every rank owns a symmetric (world, B, 2k+1) buffer (torch
    symmetric memory: each rank's allocation is mapped into every process of the node over NVLink / NVSwitch); the merge
    kernel that finishes a rank's shard stores its packed rows straight into block `rank` of ALL ranks' buffers
    (rf_cosine_topk_bcast), one symmetric-memory barrier publishes them, and rf_topk_merge_packed reads the local
    buffer.  No pack kernel, no NCCL launch: scorer -> barrier -> merge.  Two buffers alternate, so a pass may overwrite
    the buffer of the pass before last without a second barrier (every rank has passed the barrier in between, i.e. has
    merged the older one).

        ex = PeerTopkExchange(B, k, device)          # collective: every rank of the group calls it
        scores, ids, label_scores = ex.topk(model, pooled, labels=labels, id_base=lo)
    """

    def __init__(self, B: int, k: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.B, self.k = B, k
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty(self.world * B * (2 * k + 1), dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, self.group)
            self.bufs.append(t.view(self.world, B, 2 * k + 1))
            self.handles.append(h)
        self.peer_ptrs = [[int(p) for p in h.buffer_ptrs] for h in self.handles]
        self._pass = 0
        self._ws = None

    def topk(self, model, pooled: torch.Tensor, labels: Optional[torch.Tensor] = None, id_base: int = 0):
        from . import ops
        if pooled.shape[0] != self.B:
            raise ValueError(f"PeerTopkExchange was built for {self.B} users, got {pooled.shape[0]}")
        j = self._pass & 1
        self._pass += 1
        xn = ops.normalize_rows(pooled.contiguous())
        self._ws = ops.cosine_topk_bcast(xn, model.normalized_items(), model.config.temp, self.peer_ptrs[j], self.rank,
                                         k=self.k, id_base=id_base, labels=labels, ws=self._ws)
        self.handles[j].barrier(channel=j)          # on the current stream: all ranks' rows have landed everywhere
        return ops.topk_merge_packed(self.bufs[j], self.k)


def allreduce_gradients(model, group=None, average: bool = True) -> None:
    """Data-parallel gradient sync: one all-reduce over the flat fp32 gradient buffer."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    enc = getattr(model, "longformer", model)
    g = enc._engine.params.grad
    if g is None:
        raise RuntimeError("allreduce_gradients: no gradients (run backward first)")
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    if average:
        g.mul_(1.0 / world)


class GradSync:
    """Data-parallel gradient all-reduce overlapped with the backward pass.

    The engine calls `on_layer(i)` as soon as layer i's backward has been enqueued; the layer's two
    contiguous gradient ranges in the flat fp32 buffer (the six dense weights, the three *_global weights)
    are all-reduced asynchronously — NCCL orders the collective after the kernels already enqueued on the
    current stream and runs it on its own stream, next to the remaining backward kernels.  `finish()`
    reduces what is left (embedding tables, biases, LayerNorm vectors), and makes the current stream wait
    for every collective.  Gradients are SUMMED; pass `grad_scale=1/world` to `FusedAdamW.step()`.

    Wire format: on CUDA the gradients travel as bf16 (`wire_dtype`, default) — each range is cast into a flat bf16
    buffer right before its all-reduce, which halves the bytes on NVLink and the time the NCCL kernels compete with
    the backward GEMMs; the fused AdamW then reads the reduced bf16 gradients directly (`engine.params.grad_wire`).
    The reference's own data-parallel path reduces fp16 gradients (DeepSpeed stage 2, ref: lightning_pretrain.py:143).
    `wire_dtype=torch.float32` all-reduces the fp32 gradient buffer in place (CPU / gloo tests).
    """

    def __init__(self, model, group=None, passes_per_step: int = 1, wire_dtype=None, bucket_layers: int = 4):
        enc = getattr(model, "longformer", model)
        inside = {id(p) for p in enc.parameters()}
        # trainable parameters outside the encoder's flat buffer (pretraining lm_head.*): reduced in finish()
        self.extra_params = [p for p in model.parameters() if id(p) not in inside and p.requires_grad]
        self.engine = enc._engine
        self.group = group
        self.passes_per_step = passes_per_step   # encoder backward passes per optimizer step (pretraining: 4)
        self._seen = {}
        self._works = []
        self._covered = []
        self.wire_dtype = wire_dtype
        # Layers are reduced in buckets of `bucket_layers` (their dense / *_global ranges are contiguous in the flat
        # buffer): measured on 2 B200s, the 298 MB bf16 payload takes 0.61 ms as one NCCL all-reduce but 1.50 ms as 26
        # per-layer calls (launch latency + small-message efficiency), and that time is spent on SMs the backward
        # GEMMs want.  Four layers per bucket = 3 buckets + the embedding / vector tail.
        self.bucket_layers = max(1, int(bucket_layers))
        self._comm = None
        # RF_DP_SYNC_AT_END=1 (A/B aid): no per-layer overlap, one cast + one all-reduce of the whole buffer in finish()
        import os
        self.at_end = os.environ.get("RF_DP_SYNC_AT_END") is not None
        self.bucket_hook = None     # callable(ranges, works): a bucket's all-reduces were launched (FusedAdamW overlap)
        self.engine.grad_hook = self.on_layer

    def _active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def layer_ranges(self, layer: int, last_layer: Optional[int] = None):
        """The two contiguous flat-buffer ranges (dense weights, *_global weights) of layers layer..last_layer."""
        return self.engine.params.layer_ranges(layer, last_layer)

    def _wire(self, g: torch.Tensor):
        """The buffer that travels: the fp32 gradient buffer itself, or its flat bf16 twin (allocated once)."""
        dt = self.wire_dtype if self.wire_dtype is not None else (torch.bfloat16 if g.is_cuda else torch.float32)
        if dt == torch.float32:
            self.engine.params.grad_wire = None
            return None
        if self._comm is None or self._comm.numel() != g.numel() or self._comm.device != g.device:
            self._comm = torch.empty(g.numel(), dtype=torch.bfloat16, device=g.device)
        return self._comm

    def _reduce_range(self, g: torch.Tensor, a: int, b: int):
        comm = self._wire(g)
        if comm is None:
            return dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        from . import ops
        ops.cast_bf16(g[a:b], comm[a:b])
        return dist.all_reduce(comm[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def gradient(self) -> torch.Tensor:
        """The reduced (summed) flat gradient as fp32, whatever the wire format (checks / tests)."""
        P = self.engine.params
        return P.grad_wire.float() if getattr(P, "grad_wire", None) is not None else P.grad

    def on_layer(self, layer: int) -> None:
        if not self._active() or self.at_end:
            return
        self._seen[layer] = self._seen.get(layer, 0) + 1
        if self._seen[layer] < self.passes_per_step:      # gradients of this layer are still being accumulated
            return
        if layer % self.bucket_layers != 0:               # the backward visits layers top-down: a bucket is complete
            return                                        # when its lowest layer is
        n_layers = self.engine.cfg.num_hidden_layers
        g = self.engine.params.grad
        aux_done = getattr(self.engine, "aux_done", None)
        if aux_done is not None:        # weight gradients computed on the engine's aux stream: final before they travel
            torch.cuda.current_stream(g.device).wait_event(aux_done)
        ranges = self.layer_ranges(layer, min(layer + self.bucket_layers, n_layers) - 1)
        works = [self._reduce_range(g, a, b) for a, b in ranges]
        self._works += works
        self._covered += ranges
        if self.bucket_hook is not None:
            self.bucket_hook(ranges, works)

    def finish(self, defer_tail: bool = False):
        """Reduce what the per-layer calls left (embedding tables; biases / LayerNorm vectors) and wait.
        With defer_tail=True the big embedding-table range is only LAUNCHED: the returned callable waits for it,
        so the caller can run the optimizer on everything else meanwhile (FusedAdamW.step(wait_other=...))."""
        if not self._active():
            self._covered.clear()
            self._seen.clear()
            self.engine.params.grad_wire = None
            return None
        P = self.engine.params
        g = P.grad
        if g is None:
            raise RuntimeError("GradSync.finish: no gradients (run backward first)")
        gaps, pos = [], 0
        for a, b in sorted(self._covered) + [(P.n_total, P.n_total)]:
            if a > pos:
                gaps.append((pos, a))
            pos = max(pos, b)
        gaps.sort(key=lambda r: r[1] - r[0])            # small ranges first, the embedding tables last
        for p in self.extra_params:
            if p.grad is not None:
                self._works.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        tail = None
        for k, (a, b) in enumerate(gaps):
            w = self._reduce_range(g, a, b)
            if defer_tail and k == len(gaps) - 1 and P.n_dense <= a and b <= P.n_decay:
                tail = w
            else:
                self._works.append(w)
        for w in self._works:
            w.wait()
        self._works.clear()
        self._covered.clear()
        self._seen.clear()
        P.grad_wire = self._comm if self._wire(g) is not None else None     # what FusedAdamW.step() reads
        return (lambda: tail.wait()) if tail is not None else None
