"""A finetune step (forward, loss, backward, fused AdamW) captured once as a CUDA graph and replayed.

The step is ~550 kernel launches of 5-150 us each; launched one by one the GPU idles ~2 us between
consecutive kernels and stalls whenever the Python launch loop falls behind (measured with CUPTI:
~1 ms of a 14.9 ms step, tools/step_timeline.py).  Replaying the captured graph removes the host from
the loop.  What a graph freezes are kernel ARGUMENTS, so everything that changes from step to step is
read from device memory instead:
  * the batch: copied into static input tensors before each replay;
  * the dropout masks: every Philox seed is XOR-ed with a nonce the graph loads from device memory
    (rf_set_dropout_nonce), advanced before each replay;
  * AdamW's learning rate and bias corrections: rf_adamw_step_dev reads them from a device block.
The nonce and the optimiser scalars travel in one 24-byte pinned block copied to the device right
before the launch, so a replay is: fill the static inputs, copy the block, cudaGraphLaunch.  (The copy
stays outside the graph and rotates over a ring of pinned blocks: the host runs several replays ahead
of the GPU, a single block baked into the graph would be overwritten before the GPU read it.)

Data-parallel: pass the step's recformer_b200.dist.GradSync as `sync`; its per-layer asynchronous NCCL
all-reduces (and the deferred embedding-table tail) are captured with the kernels -- NCCL joins the
capture through the events torch records between the compute stream and its own.  Drop the GraphedTrainStep
(it keeps NCCL resources alive) BEFORE torch.distributed.destroy_process_group(): tearing the group down
under a live graph hangs.  Measured at 2 and 8 GPUs: 15.2 / 15.85 ms per step against 15.9 / 16.5 ms eager.
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from . import ops

_NONCE_STRIDE = 0x9E3779B97F4A7C15      # odd 64-bit constant: consecutive steps get far-apart nonces
_RING = 16


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, example_batch);  loss = step(batch)

    `model(**batch)` must return the scalar loss (RecformerForSeqRec with labels / candidates).
    `optimizer` is a recformer_b200.optim.FusedAdamW.  At least one eager step with the same batch
    shape must have run before (it sizes the engine's workspaces and the optimiser state); the
    constructor itself does not touch the parameters.  `optimizer.lr` may be changed between calls
    (schedulers): it is re-read on every replay; betas, eps, weight_decay, the dropout probabilities, the
    batch shape and which parameters are trainable are frozen at capture (build a new GraphedTrainStep after
    changing any of them).  The returned loss is a static device tensor that the next replay overwrites.
    The forward must not synchronise with the host: set `model.longformer.strict_checks = False` (the
    device-side input checks still run; `model.longformer._engine.check_errors()` reads their flags).
    """

    def __init__(self, model, optimizer, example_batch: Dict[str, torch.Tensor], grad_scale: float = 1.0, sync=None):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device")
        if optimizer.step_count < 1 or optimizer.exp_avg is None:
            raise RuntimeError("GraphedTrainStep: run one eager training step first (it allocates the engine "
                               "workspaces and the optimiser state outside the graph's memory pool)")
        self.model, self.opt, self.grad_scale = model, optimizer, grad_scale
        self.static = {k: torch.empty_like(v, device=dev) for k, v in example_batch.items()}
        for k, v in example_batch.items():
            self.static[k].copy_(v)
        # per-step block: [0:4] fp32 optimiser scalars, [4:6] the 64-bit nonce
        self._ring = [torch.zeros(6, dtype=torch.float32).pin_memory() for _ in range(_RING)]
        self._ring_f32 = [b.numpy() for b in self._ring]
        self._ring_i64 = [b[4:6].view(torch.int64).numpy() for b in self._ring]
        self._ring_done = [None] * _RING
        self._dev = torch.zeros(6, dtype=torch.float32, device=dev)
        self._dev_nonce = self._dev[4:6].view(torch.int64)
        self._replays = 0
        self.graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(dev)
        l0 = ops.launch_count()
        # Single GPU: the AdamW update of a layer is launched on the engine's aux stream as soon as that layer's backward
        # is enqueued (FusedAdamW.begin_overlap) and clears the gradients it consumed, so the step contains neither a
        # serial 0.7 ms optimiser tail nor a 600 MB memset of the gradient buffer.  RF_GRAPH_NO_OVERLAP=1 (A/B aid) or a
        # GradSync (the gradients must be all-reduced first) select the plain zero_grad / backward / step sequence.
        self.overlap = sync is None and os.environ.get("RF_GRAPH_NO_OVERLAP") is None and not optimizer.extra_params
        enc = getattr(model, "longformer", model)
        P = enc._engine.params
        untouched = optimizer.untouched_ranges() if self.overlap else []
        # Captured on a HIGH-priority stream: kernel nodes inherit it, so whenever an SM frees up the block scheduler places
        # the critical chain's CTAs before those of the aux stream (priority 0: bias-gradient column sums, overlapped
        # AdamW), which then only fill what the chain leaves.  Measured (same box, ms per step): plain 13.35, aux work at
        # equal priority 13.57 (it delays the persistent kernels' CTAs), with priorities 13.21.
        # Data-parallel steps: same priorities (NCCL's own stream, like the aux stream, has the default = lowest one: the
        # collectives fill what the backward leaves, and finish in the clear at its end); the update of a layer bucket
        # follows that bucket's all-reduce on the aux stream (FusedAdamW.begin_overlap(sync=...)).  RF_DP_PLAIN=1 (A/B
        # aid) restores the default-priority capture without aux work.
        dp_plain = sync is not None and os.environ.get("RF_DP_PLAIN") is not None
        self.dp_overlap = sync is not None and not dp_plain and os.environ.get("RF_GRAPH_NO_OVERLAP") is None \
            and not optimizer.extra_params
        prio = 0 if dp_plain else int(os.environ.get("RF_GRAPH_PRIO", "-1"))
        cap_stream = torch.cuda.Stream(device=dev, priority=prio)
        aux_before = enc._engine.overlap_aux
        enc._engine.overlap_aux = prio < 0 and os.environ.get("RF_DEBUG_NO_AUX") is None
        with torch.cuda.graph(self.graph, stream=cap_stream):
            ops.set_dropout_nonce(self._dev_nonce)
            loss = model(**self.static)
            if self.overlap:
                optimizer.begin_overlap(grad_scale=grad_scale, hp=self._dev[:4], zero_grads=True)
                for a, b in untouched:             # gradients of frozen parameters: nobody consumes (= clears) them
                    P.grad[a:b].zero_()
                loss.backward()
                optimizer.step()
            elif sync is not None and self.dp_overlap:
                optimizer.zero_grad()
                optimizer.begin_overlap(grad_scale=grad_scale, hp=self._dev[:4], sync=sync)
                loss.backward()
                optimizer.step(wait_other=sync.finish(defer_tail=True))
            else:
                optimizer.zero_grad()
                loss.backward()
                tail = sync.finish(defer_tail=True) if sync is not None else None
                optimizer.step(grad_scale=grad_scale, wait_other=tail, hp=self._dev[:4])
            self.loss = loss.detach()
        enc._engine.overlap_aux = aux_before
        if self.overlap:
            P.grad.zero_()                 # the invariant the replays keep: the gradient buffer is zero between steps
        optimizer.step_count -= 1          # capture records the launches without executing them
        self.launches_per_step = ops.launch_count() - l0

    def _send_step_block(self):
        slot = self._replays % _RING
        if self._ring_done[slot] is not None:
            self._ring_done[slot].synchronize()       # the copy that last used this pinned block has run
        self._ring_f32[slot][:4] = self.opt.step_scalars(self.grad_scale)
        self._ring_i64[slot][0] = ((self._replays + 1) * _NONCE_STRIDE) & 0x7FFFFFFFFFFFFFFF
        self._dev.copy_(self._ring[slot], non_blocking=True)
        ev = self._ring_done[slot] or torch.cuda.Event()
        ev.record()
        self._ring_done[slot] = ev

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)
        self._send_step_block()
        self._replays += 1
        self.graph.replay()
        self.opt.step_count += 1
        return self.loss
