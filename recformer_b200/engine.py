"""Host-side engine of the B200 encoder: flat parameter storage, activation workspaces and the
forward / backward launch schedules over the C-ABI kernels.

Data layout in HBM
  * parameters live in ONE flat fp32 buffer (the nn.Parameters of the reference-compatible module
    tree are views into it), ordered so that query/key/value weights of a layer are contiguous
    (= the fused [3E,E] QKV weight) and so that AdamW can run as two launches (decay / no-decay);
  * the dense encoder weights additionally have a bf16 shadow in the same order (TMA/tcgen05
    operands); gradients live in a flat fp32 buffer with the same offsets, `.grad`s are views;
  * the residual stream (pre-LayerNorm sums and LayerNorm outputs) is fp32 [B*Lp, E]; every
    tensor that is a tensor-core operand (LN outputs, qkv, attention context, GELU in/out) is
    bf16 [B*Lp, *] row-major; statistics (LN mean/rstd, attention LSE) are fp32.  This is the
    numerics of bf16 autocast (measured need: with a bf16 residual stream the C1 logits drift
    2.9e-2 from the fp32 reference, above the 2e-2 budget).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import ops


def _pick_split(M: int, N: int, K: int, sms: int = 148) -> int:
    """Split-K factor for a weight-gradient GEMM (K = tokens): fill the 74 CTA pairs that each own
    a 256 x 256 output tile, keeping at least 8 K-blocks (512 tokens) per slice."""
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    pairs = sms // 2
    kb = (K + 63) // 64
    best, best_eff = 1, 0.0
    for s in range(1, 17):
        if s > 1 and kb // s < 8:
            break
        total = tiles * s
        eff = total / (((total + pairs - 1) // pairs) * pairs)
        if eff > best_eff + 0.02:
            best, best_eff = s, eff
    return best


class FlatParams:
    """Flat fp32 parameter / gradient buffers + bf16 shadow of the dense encoder weights."""

    def __init__(self, model):
        self.model = model
        self.flat: Optional[torch.Tensor] = None
        self.grad: Optional[torch.Tensor] = None
        self.grad_wire: Optional[torch.Tensor] = None   # bf16 twin holding the all-reduced gradients (dist.GradSync)
        self.shadow: Optional[torch.Tensor] = None
        self._optimizer_fresh = False
        self.offsets: Dict[str, int] = {}
        self.n_dense = 0
        self.n_decay = 0
        self.n_total = 0
        self._order: List[str] = []
        self._sentinel = None
        self._shadow_sig = None
        self._plan()

    def _plan(self):
        m = self.model
        named = dict(m.named_parameters())
        dense, other, nodecay = [], [], []
        nl = m.config.num_hidden_layers
        for i in range(nl):
            p = f"encoder.layer.{i}."
            dense += [p + "attention.self.query.weight", p + "attention.self.key.weight",
                      p + "attention.self.value.weight", p + "attention.output.dense.weight",
                      p + "intermediate.dense.weight", p + "output.dense.weight"]
            other += [p + "attention.self.query_global.weight", p + "attention.self.key_global.weight",
                      p + "attention.self.value_global.weight"]
            nodecay += [p + "attention.self.query.bias", p + "attention.self.key.bias", p + "attention.self.value.bias",
                        p + "attention.self.query_global.bias", p + "attention.self.key_global.bias",
                        p + "attention.self.value_global.bias", p + "attention.output.dense.bias",
                        p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
                        p + "intermediate.dense.bias", p + "output.dense.bias",
                        p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"]
        other = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
                 "embeddings.token_type_embeddings.weight", "embeddings.item_position_embeddings.weight"] + other
        nodecay += ["embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]
        order = dense + other + nodecay
        assert set(order) == set(named), (set(named) ^ set(order))
        off = 0
        for k in order:
            n = named[k].numel()
            assert n % 8 == 0, (k, n)
            self.offsets[k] = off
            off += n
            if k == dense[-1]:
                self.n_dense = off
            if k == other[-1]:
                self.n_decay = off
        self.n_total = off
        self._order = order
        self._named = named

    # -- storage -------------------------------------------------------------------------------
    def _aliased(self, device) -> bool:
        """True when every Parameter object of the module tree is still the one planned AND still a view at its
        offset of the flat buffer (a replaced module / Parameter, load_state_dict(assign=True) or .to() breaks it)."""
        if self.flat is None or self.flat.device != torch.device(device):
            return False
        base = self.flat.data_ptr()
        live = dict(self.model.named_parameters())
        if len(live) != len(self._order):
            return False
        for k in self._order:
            p = live.get(k)
            if p is None or p is not self._named[k] or p.data_ptr() != base + 4 * self.offsets[k]:
                return False
        return True

    def ensure(self, device) -> None:
        if self._aliased(device):
            return
        live = dict(self.model.named_parameters())
        if set(live) != set(self._order) or any(live[k].shape != self._named[k].shape for k in self._order):
            raise RuntimeError("recformer_b200: the module tree's parameters changed shape or name since the engine was "
                               "planned; rebuild the model")
        self._named = live                      # adopt replaced Parameter objects (their values are copied below)
        first = self._named[self._order[0]]
        if first.device != torch.device(device):
            raise RuntimeError(f"recformer_b200: parameters are on {first.device}, inputs on {device}")
        if first.device.type != "cuda":
            raise RuntimeError("recformer_b200 runs on CUDA only: move the model with .cuda() / .to('cuda') "
                               "(there is no CPU path)")
        flat = torch.empty(self.n_total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for k in self._order:
                p = self._named[k]
                o = self.offsets[k]
                view = flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().to(torch.float32))
                p.data = view
                p.grad = None
        self.flat = flat
        self.grad = None
        self.shadow = torch.empty(self.n_dense, dtype=torch.bfloat16, device=device)
        self._shadow_sig = None

    def view(self, key: str, flat: Optional[torch.Tensor] = None, shape=None) -> torch.Tensor:
        p = self._named[key]
        o = self.offsets[key]
        base = self.flat if flat is None else flat
        return base[o:o + p.numel()].view(p.shape if shape is None else shape)

    def fused(self, first_key: str, n_params: int, base: torch.Tensor, shape) -> torch.Tensor:
        """View spanning `n_params` consecutive parameters starting at `first_key` (fused QKV)."""
        o = self.offsets[first_key]
        n = 1
        for s in shape:
            n *= s
        return base[o:o + n].view(shape)

    def layer_ranges(self, layer: int, last_layer: Optional[int] = None):
        """The two contiguous flat-buffer ranges (six dense weights, three *_global weights) of layers layer..last_layer."""
        p, q = f"encoder.layer.{layer}.", f"encoder.layer.{layer if last_layer is None else last_layer}."
        named = self._named
        d0 = self.offsets[p + "attention.self.query.weight"]
        d1 = self.offsets[q + "output.dense.weight"] + named[q + "output.dense.weight"].numel()
        g0 = self.offsets[p + "attention.self.query_global.weight"]
        g1 = self.offsets[q + "attention.self.value_global.weight"] + named[q + "attention.self.value_global.weight"].numel()
        return [(d0, d1), (g0, g1)]

    def invalidate(self) -> None:
        """Force the next forward to recast the bf16 shadow (after weights were edited through `.data`, which
        does not bump the autograd version counters refresh_shadow() keys on)."""
        self._shadow_sig = None
        self._optimizer_fresh = False

    def refresh_shadow(self, force: bool = False) -> None:
        sig = sum(self._named[k]._version for k in self._order[: 6 * self.model.config.num_hidden_layers])
        if self._optimizer_fresh and sig == self._shadow_sig:
            # the fused AdamW wrote the bf16 shadow together with the fp32 weights: nothing to cast, once
            self._optimizer_fresh = False
            return
        self._optimizer_fresh = False
        if force or sig != self._shadow_sig:
            ops.cast_bf16(self.flat[: self.n_dense], self.shadow)
            self._shadow_sig = sig

    def mark_shadow_fresh(self, by_optimizer: bool = False) -> None:
        self._shadow_sig = sum(self._named[k]._version for k in self._order[: 6 * self.model.config.num_hidden_layers])
        self._optimizer_fresh = by_optimizer

    def prepare_grads(self) -> None:
        """Make every trainable parameter's .grad a view of the flat gradient buffer.  If no
        parameter currently has a gradient (optimizer.zero_grad(set_to_none=True)) the whole
        buffer is zeroed with one memset; existing view-gradients are accumulated into."""
        if self.grad is None:
            self.grad = torch.zeros(self.n_total, dtype=torch.float32, device=self.flat.device)
            fresh = True
        else:
            fresh = False
        params = [(k, self._named[k]) for k in self._order]
        if all(p.grad is None for _, p in params):
            if not fresh:
                self.grad.zero_()
            for k, p in params:
                if p.requires_grad:
                    p.grad = self.view(k, self.grad)
            return
        for k, p in params:
            if not p.requires_grad:
                continue
            gv = self.view(k, self.grad)
            if p.grad is None:
                gv.zero_()
                p.grad = gv
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)      # foreign gradient tensor: adopt its value, then accumulate in place
                p.grad = gv


# A/B aid: RF_NO_KEEPBITS=1 makes every backward kernel regenerate its dropout mask instead of reading the saved bits
_SAVE_KEEPBITS = os.environ.get("RF_NO_KEEPBITS", "0") != "1"


_GLOBAL_WGRAD_INLINE = os.environ.get("RF_GLOBAL_WGRAD_INLINE", "0") == "1"     # A/B aid: the former launch order


class SavedActivations:
    """Everything one forward pass keeps for its backward (or, in eval, reusable scratch)."""

    def __init__(self, B: int, Lp: int, cfg, device, per_layer: bool):
        E, F, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
        T = B * Lp
        nl = cfg.num_hidden_layers
        n = nl if per_layer else 1
        bf = dict(dtype=torch.bfloat16, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        self.B, self.Lp, self.per_layer = B, Lp, per_layer
        self.x = [torch.zeros(T, E, **bf) for _ in range(nl + 1 if per_layer else 2)]      # bf16 operand copy
        self.x32 = [torch.zeros(T, E, **f32) for _ in range(2)]                             # fp32 residual stream
        self.h1_32 = torch.zeros(T, E, **f32)
        self.qkv = [torch.zeros(T, 3 * E, **bf) for _ in range(n)]
        self.lse = [torch.zeros(B, H, Lp, **f32) for _ in range(n)]
        self.ctx = [torch.zeros(T, E, **bf) for _ in range(n)]
        # attention-probability dropout keep bits (16 B per row and head), saved by the forward for the backward
        self.keepbits = [ops.band_attn_keepbits(B, Lp, H, device) if per_layer and _SAVE_KEEPBITS else None for _ in range(n)]
        # hidden-dropout masks of the two residual GEMMs (one byte per 8 columns), saved for the LayerNorm backward
        self.mask1 = [torch.zeros(T, E // 8, dtype=torch.uint8, device=device) if per_layer and _SAVE_KEEPBITS else None
                      for _ in range(n)]
        self.mask2 = [torch.zeros(T, E // 8, dtype=torch.uint8, device=device) if per_layer and _SAVE_KEEPBITS else None
                      for _ in range(n)]
        self.pre1 = [torch.zeros(T, E, **f32) for _ in range(n)]
        self.stats1 = [torch.zeros(T, 2, **f32) for _ in range(n)]
        self.h1 = [torch.zeros(T, E, **bf) for _ in range(n)]
        self.u = [torch.zeros(T, F, **bf) for _ in range(n)]       # gelu'(pre-activation), saved for backward
        self.g = [torch.zeros(T, F, **bf) for _ in range(n)]
        self.pre2 = [torch.zeros(T, E, **f32) for _ in range(n)]
        self.stats2 = [torch.zeros(T, 2, **f32) for _ in range(n)]
        self.glob = [ops.global_attn_saved(B, Lp, H, device) for _ in range(n)]
        # padding-aware execution (include/recformer_b200.h, rf_set_row_activity): 256-row tiles that hold a real token.
        # Buffers above are zero-initialised because rows of skipped tiles are never written: whatever a kernel reads
        # from them (keys of a masked neighbour tile under a zero probability) must at least be finite.
        self.activity, self.skip = None, False
        if Lp % 256 == 0:
            self.activity = (torch.zeros(T // 256, dtype=torch.uint8, device=device),
                             torch.zeros(T // 128, dtype=torch.int32, device=device),
                             torch.zeros(1, dtype=torch.int32, device=device))
        self.pos_ids = None
        self.mask012 = None
        self.inputs = None
        self.seed = 0
        self.drop_hidden = 0.0
        self.drop_attn = 0.0

    def idx(self, layer: int) -> int:
        return layer if self.per_layer else 0

    def xin(self, layer: int) -> torch.Tensor:
        return self.x[layer] if self.per_layer else self.x[layer % 2]

    def xout(self, layer: int) -> torch.Tensor:
        return self.x[layer + 1] if self.per_layer else self.x[(layer + 1) % 2]


class BackwardScratch:
    def __init__(self, B: int, Lp: int, cfg, device):
        E, F, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
        T = B * Lp
        bf = dict(dtype=torch.bfloat16, device=device)
        # Everything the aux stream reads (weight-gradient GEMMs, bias column sums) exists once per LayerNorm site and
        # layer parity: the main stream may be up to two layers ahead of the aux stream before it has to wait.
        self.d_pre = [[torch.zeros(T, E, **bf) for _ in range(2)] for _ in range(2)]         # [site][parity]
        self.d_pre_drop = [[torch.zeros(T, E, **bf) for _ in range(2)] for _ in range(2)]
        self.dU = [torch.zeros(T, F, **bf) for _ in range(2)]
        self.dh1 = torch.zeros(T, E, **bf)
        self.dctx = torch.zeros(T, E, **bf)
        self.dqkv = [torch.zeros(T, 3 * E, **bf) for _ in range(2)]
        self.dx = [torch.zeros(T, E, **bf) for _ in range(2)]
        self.dkv = torch.zeros(T, 2 * E, dtype=torch.float32, device=device)
        self.gws = ops.global_attn_bwd_ws(B, Lp, H, device)
        # operands of the rank-64 per-sequence update that carries the CLS row's token gradients through the
        # QKV dgrad GEMM (only when a 256-row output tile never straddles two sequences)
        self.xk = (torch.zeros(T, 64, **bf), torch.zeros(B * 64, E, **bf)) if Lp % 256 == 0 and T > 128 else None


class EncoderEngine:
    def __init__(self, model):
        self.model = model
        self.cfg = model.config
        self.params = FlatParams(model)
        self._free: Dict[tuple, List[SavedActivations]] = {}
        self._bwd: Dict[tuple, BackwardScratch] = {}
        self._attn_ws: Dict[tuple, Optional[torch.Tensor]] = {}     # wide-window scratch, never freed (graphs hold it)
        self._dout_cls: Dict[tuple, torch.Tensor] = {}              # all-zero upstream-gradient buffers (CLS-only backward)
        self._err = None
        self._call = 0
        # The global (CLS) row only depends on the layer input (forward) / on dctx (backward) and is made
        # of small latency-bound kernels: it runs on a side stream next to the QKV GEMM + band attention.
        self._side: Dict[str, torch.cuda.Stream] = {}
        self._events: Dict[tuple, torch.cuda.Event] = {}
        self.overlap_global = os.environ.get("RF_DEBUG_NO_OVERLAP") is None
        self._debug_skip_global = os.environ.get("RF_DEBUG_SKIP_GLOBAL") is not None   # timing experiment only (wrong results)
        self._debug_no_xk = os.environ.get("RF_DEBUG_NO_XK") is not None   # A/B switch: separate dx-update kernel
        self.tile_skip = os.environ.get("RF_NO_TILE_SKIP") is None     # padding-aware execution (forward(skip_padding=True))
        self.grad_hook = None      # callable(layer): that layer's gradients are final (dist.GradSync, FusedAdamW overlap)
        # Second side stream ("aux"): work nobody on the backward's critical path waits for — the bias-gradient column
        # sums of dU / dqkv and, when FusedAdamW.begin_overlap() armed it, the AdamW update of finished layers.  These
        # are small-footprint HBM-bound kernels (<= 40 registers, <= 8 KB of shared memory) that fit on an SM NEXT to a
        # resident persistent GEMM CTA, so they use the HBM bandwidth the tensor-bound GEMMs leave idle.
        self._aux: Dict[str, torch.cuda.Stream] = {}
        # Off by default: on a stream of the SAME priority as the main one this work delays the CTAs of the persistent
        # kernels (13.57 vs 13.35 ms per step); GraphedTrainStep captures on a high-priority stream and switches it on.
        self.overlap_aux = False
        # With overlap_aux the four weight-gradient GEMMs of a layer also go to the aux stream: nothing on the backward's
        # dependency chain reads them, and their (dynamically scheduled) CTAs fill the last, partly empty wave of the
        # chain's GEMMs.  Measured with every row tile active: no gain (13.17-13.35 vs 13.23-13.30 ms; the step runs at
        # the board's power cap and busier SMs are paid back in clock, 1725 vs 1800 MHz); with padding tiles skipped the
        # N = 768 GEMMs have 144-192 tiles on 74 CTA pairs, i.e. long tails: 12.72-12.79 vs 12.91-12.97 ms.
        # RF_NO_WGRAD_AUX=1 keeps them on the main stream.
        self.wgrad_aux = os.environ.get("RF_NO_WGRAD_AUX") is None
        self.aux_done = None       # latest aux-stream event of the running backward (dist.GradSync waits on it)

    # -- helpers -------------------------------------------------------------------------------
    def _acquire(self, B, Lp, device, per_layer) -> SavedActivations:
        key = (B, Lp, str(device), per_layer)
        pool = self._free.setdefault(key, [])
        return pool.pop() if pool else SavedActivations(B, Lp, self.cfg, device, per_layer)

    def release(self, sv: SavedActivations) -> None:
        key = (sv.B, sv.Lp, str(sv.x[0].device), sv.per_layer)
        pool = self._free.setdefault(key, [])
        if len(pool) < 4:
            pool.append(sv)

    def attn_ws(self, B: int, Lp: int, w_one: int, device) -> Optional[torch.Tensor]:
        """Engine-owned band-attention scratch for windows wider than 64, one per (B, Lp, w) and kept for the
        engine's life: a captured step graph holds its address, and no other engine or stream ever sees it."""
        key = (B, Lp, w_one, str(device))
        if key not in self._attn_ws:
            self._attn_ws[key] = ops.band_attn_ws(B, Lp, self.cfg.num_attention_heads, w_one, device)
        return self._attn_ws[key]

    def side_stream(self, device) -> torch.cuda.Stream:
        key = str(device)
        if key not in self._side:
            # high priority: the main stream waits for this chain of small kernels once per layer and direction
            self._side[key] = torch.cuda.Stream(device=device, priority=int(os.environ.get("RF_SIDE_PRIO", "-1")))
        return self._side[key]

    def aux_stream(self, device) -> torch.cuda.Stream:
        key = str(device)
        if key not in self._aux:
            self._aux[key] = torch.cuda.Stream(device=device, priority=0)      # lowest: runs in what the others leave
        return self._aux[key]

    def fork_aux(self, device, name: str, layer: int, fn) -> torch.cuda.Event:
        """Run fn() on the aux stream after everything enqueued so far on the current stream; returns the event that
        marks its completion (the caller waits on it before it overwrites what fn reads, and at the end of the pass)."""
        main, aux = torch.cuda.current_stream(device), self.aux_stream(device)
        ev = self.event(device, name, layer)
        ev.record(main)
        aux.wait_event(ev)
        with torch.cuda.stream(aux):
            fn()
            done = self.event(device, name + ".done", layer)
            done.record(aux)
        return done

    def event(self, device, name: str, layer: int) -> torch.cuda.Event:
        key = (str(device), name, layer)
        ev = self._events.get(key)
        if ev is None:
            ev = self._events[key] = torch.cuda.Event()
        return ev

    def err_flag(self, device) -> torch.Tensor:
        if self._err is None or self._err.device != torch.device(device):
            self._err = torch.zeros(1, dtype=torch.int32, device=device)
        return self._err

    def check_errors(self) -> None:
        """Synchronising check of the device-side input validation flags."""
        if self._err is None:
            return
        v = int(self._err.item())
        if v:
            self._err.zero_()
            msgs = []
            if v & 1:
                msgs.append("global_attention_mask marks a position other than 0 (only the tokenizer's CLS-global "
                            "layout is supported, ref: recformer/tokenization.py:97-99)")
            if v & 2:
                msgs.append("an input / token-type / item-position / position id is out of range")
            raise ValueError("recformer_b200: " + "; ".join(msgs))

    def window_pad(self, L: int) -> int:
        aw = self.cfg.attention_window
        w = aw if isinstance(aw, int) else max(aw)
        return (L + w - 1) // w * w

    def _layer_weights(self, i: int):
        P = self.params
        E, F = self.cfg.hidden_size, self.cfg.intermediate_size
        p = f"encoder.layer.{i}."
        sh, fl = P.shadow, P.flat
        return {
            "Wqkv": P.fused(p + "attention.self.query.weight", 3, sh, (3 * E, E)),
            "bqkv": P.fused(p + "attention.self.query.bias", 3, fl, (3 * E,)),
            "Wo": P.view(p + "attention.output.dense.weight", sh), "bo": P.view(p + "attention.output.dense.bias"),
            "W1": P.view(p + "intermediate.dense.weight", sh), "b1": P.view(p + "intermediate.dense.bias"),
            "W2": P.view(p + "output.dense.weight", sh), "b2": P.view(p + "output.dense.bias"),
            "ln1w": P.view(p + "attention.output.LayerNorm.weight"), "ln1b": P.view(p + "attention.output.LayerNorm.bias"),
            "ln2w": P.view(p + "output.LayerNorm.weight"), "ln2b": P.view(p + "output.LayerNorm.bias"),
            "Wqg": P.view(p + "attention.self.query_global.weight"), "bqg": P.view(p + "attention.self.query_global.bias"),
            "Wkg": P.view(p + "attention.self.key_global.weight"),
            "Wvg": P.view(p + "attention.self.value_global.weight"), "bvg": P.view(p + "attention.self.value_global.bias"),
        }

    def _layer_grads(self, i: int):
        P = self.params
        E = self.cfg.hidden_size
        p = f"encoder.layer.{i}."
        g = P.grad
        return {
            "Wqkv": P.fused(p + "attention.self.query.weight", 3, g, (3 * E, E)),
            "bqkv": P.fused(p + "attention.self.query.bias", 3, g, (3 * E,)),
            "Wo": P.view(p + "attention.output.dense.weight", g), "bo": P.view(p + "attention.output.dense.bias", g),
            "W1": P.view(p + "intermediate.dense.weight", g), "b1": P.view(p + "intermediate.dense.bias", g),
            "W2": P.view(p + "output.dense.weight", g), "b2": P.view(p + "output.dense.bias", g),
            "ln1w": P.view(p + "attention.output.LayerNorm.weight", g), "ln1b": P.view(p + "attention.output.LayerNorm.bias", g),
            "ln2w": P.view(p + "output.LayerNorm.weight", g), "ln2b": P.view(p + "output.LayerNorm.bias", g),
            "Wqg": P.view(p + "attention.self.query_global.weight", g), "bqg": P.view(p + "attention.self.query_global.bias", g),
            "Wkg": P.view(p + "attention.self.key_global.weight", g),
            "Wvg": P.view(p + "attention.self.value_global.weight", g), "bvg": P.view(p + "attention.self.value_global.bias", g),
        }

    def _seed(self, sv: SavedActivations, layer: int, site: int) -> int:
        return (sv.seed + 1000003 * (layer + 1) + 7919 * site) & 0x7FFFFFFFFFFFFFFF

    # -- forward -------------------------------------------------------------------------------
    def forward(self, input_ids, attention_mask, global_attention_mask, token_type_ids, item_position_ids,
                position_ids=None, training: bool = False, save: bool = False,
                skip_padding: bool = False) -> SavedActivations:
        """skip_padding: the caller only consumes rows of real tokens (the CLS rows, gathered masked rows) and feeds
        back zero gradient for padded positions: 256-row tiles made of padding only are skipped by every token-major
        kernel of the pass and of its backward (rf_set_row_activity); their rows of the hidden states keep whatever an
        earlier pass left there.  Needs Lp % 256 == 0, otherwise ignored."""
        cfg, P = self.cfg, self.params
        device = input_ids.device
        P.ensure(device)
        P.refresh_shadow(force=training and save)
        B, L = input_ids.shape
        Lp = self.window_pad(L)
        E, H = cfg.hidden_size, cfg.num_attention_heads
        sv = self._acquire(B, Lp, device, save)
        sv.drop_hidden = float(cfg.hidden_dropout_prob) if training else 0.0
        sv.drop_attn = float(cfg.attention_probs_dropout_prob) if training else 0.0
        self._call += 1
        sv.seed = (torch.initial_seed() * 2654435761 + self._call * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF
        err = self.err_flag(device)
        pos, mask = ops.prepare_inputs(input_ids, attention_mask, global_attention_mask, Lp, cfg.pad_token_id, err)
        if position_ids is not None:
            pos = torch.nn.functional.pad(position_ids.to(torch.int32), (0, Lp - L), value=cfg.pad_token_id).contiguous()
        sv.pos_ids, sv.mask012 = pos, mask
        sv.inputs = (input_ids, token_type_ids, item_position_ids)
        sv.skip = bool(skip_padding) and self.tile_skip and sv.activity is not None
        if sv.skip:
            ops.row_tile_flags(mask, B, Lp, *sv.activity)
            ops.set_row_activity(sv.activity[0], B * Lp, sv.activity[1], sv.activity[2])
        try:
            self._forward_layers(sv, input_ids, token_type_ids, item_position_ids, pos, mask, err, device)
        finally:
            ops.set_row_activity(None)
        return sv

    def _forward_layers(self, sv, input_ids, token_type_ids, item_position_ids, pos, mask, err, device) -> None:
        cfg, P = self.cfg, self.params
        B, Lp = sv.B, sv.Lp
        E, H = cfg.hidden_size, cfg.num_attention_heads
        e = "embeddings."
        ops.embed_ln_fwd(input_ids, token_type_ids, item_position_ids, pos,
                         P.view(e + "word_embeddings.weight"), P.view(e + "position_embeddings.weight"),
                         P.view(e + "token_type_embeddings.weight"), P.view(e + "item_position_embeddings.weight"),
                         P.view(e + "LayerNorm.weight"), P.view(e + "LayerNorm.bias"), Lp, cfg.pad_token_id,
                         cfg.layer_norm_eps, err, drop_p=sv.drop_hidden, drop_seed=self._seed(sv, -1, 0),
                         out=sv.xin(0), out32=sv.x32[0])
        aw = cfg.attention_window
        for i in range(cfg.num_hidden_layers):
            W = self._layer_weights(i)
            k = sv.idx(i)
            x = sv.xin(i)
            w_one = (aw if isinstance(aw, int) else aw[i]) // 2
            if self._debug_skip_global:
                pass
            elif self.overlap_global:
                # global row (writes ctx row 0 of every sequence) on the side stream, concurrently with the
                # QKV projection + band attention (which never touch that row when position 0 is global)
                main, side = torch.cuda.current_stream(device), self.side_stream(device)
                ev_x, ev_g = self.event(device, "x", i), self.event(device, "g", i)
                ev_x.record(main)
                side.wait_event(ev_x)
                with torch.cuda.stream(side):
                    ops.global_attn_fwd(x, mask, W["Wqg"], W["bqg"], W["Wkg"], W["Wvg"], W["bvg"], B, Lp, H, sv.ctx[k],
                                        saved=sv.glob[k], drop_p=sv.drop_attn, drop_seed=self._seed(sv, i, 2))
                    ev_g.record(side)
            ops.gemm(x, W["Wqkv"], out=sv.qkv[k], bias=W["bqkv"], scale=0.125, scale_ncols=E)
            ops.band_attn_fwd(sv.qkv[k], mask, B, Lp, H, w_one, ctx=sv.ctx[k], lse=sv.lse[k], drop_p=sv.drop_attn,
                              drop_seed=self._seed(sv, i, 1), ws=self.attn_ws(B, Lp, w_one, device),
                              keepbits=sv.keepbits[k])
            if self._debug_skip_global:
                pass
            elif self.overlap_global:
                main.wait_event(ev_g)
            else:
                ops.global_attn_fwd(x, mask, W["Wqg"], W["bqg"], W["Wkg"], W["Wvg"], W["bvg"], B, Lp, H, sv.ctx[k],
                                    saved=sv.glob[k], drop_p=sv.drop_attn, drop_seed=self._seed(sv, i, 2))
            ops.gemm(sv.ctx[k], W["Wo"], out=sv.pre1[k], bias=W["bo"], residual=sv.x32[i % 2],
                     drop_p=sv.drop_hidden, drop_seed=self._seed(sv, i, 3), drop_mask=sv.mask1[k])
            ops.layernorm_fwd(sv.pre1[k], W["ln1w"], W["ln1b"], cfg.layer_norm_eps, out=sv.h1[k], out32=sv.h1_32,
                              stats=sv.stats1[k])
            ops.gemm(sv.h1[k], W["W1"], out=sv.u[k], bias=W["b1"], epi=ops.EPI_GELU, out2=sv.g[k])
            ops.gemm(sv.g[k], W["W2"], out=sv.pre2[k], bias=W["b2"], residual=sv.h1_32, drop_p=sv.drop_hidden,
                     drop_seed=self._seed(sv, i, 4), drop_mask=sv.mask2[k])
            ops.layernorm_fwd(sv.pre2[k], W["ln2w"], W["ln2b"], cfg.layer_norm_eps, out=sv.xout(i),
                              out32=sv.x32[(i + 1) % 2], stats=sv.stats2[k])

    def hidden(self, sv: SavedActivations) -> torch.Tensor:
        """fp32 [B*Lp, E] final hidden states (the residual-stream copy of the last LayerNorm)."""
        return sv.x32[self.cfg.num_hidden_layers % 2]

    def hidden_bf16(self, sv: SavedActivations) -> torch.Tensor:
        nl = self.cfg.num_hidden_layers
        return sv.x[nl] if sv.per_layer else sv.x[nl % 2]

    # -- backward ------------------------------------------------------------------------------
    def backward_from_pooled(self, sv: SavedActivations, d_pooled: torch.Tensor) -> None:
        """Backward when only the CLS rows carry gradient (RecformerPooler 'cls', ref: recformer/models.py:160-171):
        d_pooled [B, E].  The [B*Lp, E] bf16 upstream gradient is a persistent all-zero buffer of the engine whose B CLS
        rows are filled in and cleared again afterwards, instead of a fresh 25 MB memset + strided copy per step."""
        key = (sv.B, sv.Lp, str(d_pooled.device))
        buf = self._dout_cls.get(key)
        if buf is None:
            buf = self._dout_cls[key] = torch.zeros(sv.B * sv.Lp, self.cfg.hidden_size, dtype=torch.bfloat16,
                                                    device=d_pooled.device)
        rows = buf.view(sv.B, sv.Lp, -1)[:, 0]
        rows.copy_(d_pooled)
        self.backward(sv, buf)
        rows.zero_()

    def backward(self, sv: SavedActivations, dout: torch.Tensor) -> None:
        """dout: bf16 [B*Lp, E] gradient w.r.t. the final hidden states.  Accumulates into .grad."""
        if sv.skip:
            ops.set_row_activity(sv.activity[0], sv.B * sv.Lp, sv.activity[1], sv.activity[2])
        try:
            self._backward(sv, dout)
        finally:
            ops.set_row_activity(None)

    def _backward(self, sv: SavedActivations, dout: torch.Tensor) -> None:
        cfg, P = self.cfg, self.params
        assert sv.per_layer, "forward was not run with save=True"
        B, Lp = sv.B, sv.Lp
        E, F, H = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
        T = B * Lp
        device = dout.device
        P.prepare_grads()
        key = (B, Lp, str(device))
        sc = self._bwd.get(key)
        if sc is None:
            sc = self._bwd[key] = BackwardScratch(B, Lp, cfg, device)
        mask = sv.mask012
        pd = sv.drop_hidden
        aw = cfg.attention_window
        d_out = dout
        main = torch.cuda.current_stream(device)
        use_aux = self.overlap_aux
        wg_aux = use_aux and self.wgrad_aux
        aux_marks = {}                  # layer -> aux-stream event after that layer's last aux-stream kernel
        self.aux_done = None

        def on_aux(name, layer, fn):
            self.aux_done = self.fork_aux(device, name, layer, fn)

        for i in reversed(range(cfg.num_hidden_layers)):
            W, G = self._layer_weights(i), self._layer_grads(i)
            x = sv.x[i]
            w_one = (aw if isinstance(aw, int) else aw[i]) // 2
            par = i & 1
            if aux_marks.get(i + 2) is not None:     # this parity's scratch was last read by the aux stream two layers up
                main.wait_event(aux_marks[i + 2])
            d_pre2, d_pre1 = sc.d_pre[0][par], sc.d_pre[1][par]
            dY2 = sc.d_pre_drop[0][par] if pd > 0 else d_pre2
            dY1 = sc.d_pre_drop[1][par] if pd > 0 else d_pre1
            dU, dqkv = sc.dU[par], sc.dqkv[par]

            def wgrad(name, dy, act, out, M, N):
                fn = lambda: ops.gemm(dy, act, out=out, a_mn_major=True, b_mn_major=True, accumulate=True,
                                      split_k=_pick_split(M, N, T))
                if wg_aux:
                    on_aux(name, i, fn)
                else:
                    fn()

            ev_w = None
            # ---- output block: LN2 <- dense(W2) <- gelu <- dense(W1) ----
            ops.layernorm_bwd(d_out, sv.pre2[i], sv.stats2[i], W["ln2w"], G["ln2w"], G["ln2b"], dx=d_pre2,
                              dx_dropped=dY2 if pd > 0 else None, drop_p=pd, drop_seed=self._seed(sv, i, 4),
                              d_bias=G["b2"], drop_mask=sv.mask2[i])
            wgrad("wgW2", dY2, sv.g[i], G["W2"], E, F)
            ops.gemm(dY2, W["W2"], out=dU, b_mn_major=True, epi=ops.EPI_DGELU, aux=sv.u[i])
            if use_aux:
                on_aux("dU", i, lambda: ops.colsum(dU, G["b1"]))
            else:
                ops.colsum(dU, G["b1"])
            wgrad("wgW1", dU, sv.h1[i], G["W1"], F, E)
            ops.gemm(dU, W["W1"], out=sc.dh1, b_mn_major=True, residual=d_pre2)
            # ---- attention block: LN1 <- dense(Wo) <- attention <- dense(Wqkv) ----
            ops.layernorm_bwd(sc.dh1, sv.pre1[i], sv.stats1[i], W["ln1w"], G["ln1w"], G["ln1b"], dx=d_pre1,
                              dx_dropped=dY1 if pd > 0 else None, drop_p=pd, drop_seed=self._seed(sv, i, 3),
                              d_bias=G["bo"], drop_mask=sv.mask1[i])
            wgrad("wgWo", dY1, sv.ctx[i], G["Wo"], E, E)
            ops.gemm(dY1, W["Wo"], out=sc.dctx, b_mn_major=True)
            gargs = (x, mask, W["Wqg"], W["bqg"], W["Wkg"], W["Wvg"], W["bvg"], B, Lp, H)
            if self._debug_skip_global:
                pass
            elif self.overlap_global:
                # weight-gradient half of the global row's backward (needs dctx, not dx) on the side stream,
                # concurrently with the band-attention backward and the QKV wgrad / dgrad GEMMs
                side = self.side_stream(device)
                ev_d, ev_a = self.event(device, "dctx", i), self.event(device, "gA", i)
                ev_d.record(main)
                side.wait_event(ev_d)
                with torch.cuda.stream(side):
                    # the three *_global weight gradients feed only the optimiser: they are launched AFTER the operands
                    # the main stream waits for (ev_a), off the critical chain; ev_w joins them at the end of the layer
                    late = not _GLOBAL_WGRAD_INLINE
                    ops.global_attn_bwd(*gargs, sc.dctx, sv.glob[i], None, None if late else G["Wqg"], G["bqg"],
                                        None if late else G["Wkg"], None if late else G["Wvg"],
                                        G["bvg"], ws=sc.gws, drop_p=sv.drop_attn, drop_seed=self._seed(sv, i, 2))
                    if sc.xk is not None and not self._debug_no_xk:
                        ops.global_attn_bwd_xk(*gargs, sv.glob[i], sc.gws, *sc.xk)
                    ev_a.record(side)
                    if late:
                        ops.global_attn_bwd_wgrad(*gargs, sv.glob[i], sc.gws, G["Wqg"], G["Wkg"], G["Wvg"])
                        ev_w = self.event(device, "gW", i)
                        ev_w.record(side)
            ops.band_attn_bwd(sv.qkv[i], mask, B, Lp, H, w_one, sv.ctx[i], sv.lse[i], sc.dctx, dqkv, sc.dkv,
                              drop_p=sv.drop_attn, drop_seed=self._seed(sv, i, 1), ws=self.attn_ws(B, Lp, w_one, device),
                              keepbits=sv.keepbits[i])
            if use_aux:
                on_aux("dqkv", i, lambda: ops.colsum(dqkv, G["bqkv"]))
            else:
                ops.colsum(dqkv, G["bqkv"])
            wgrad("wgWqkv", dqkv, x, G["Wqkv"], 3 * E, E)
            dx = sc.dx[i % 2]
            fused_dx = self.overlap_global and sc.xk is not None and not self._debug_skip_global and not self._debug_no_xk
            if fused_dx:
                # the CLS row's token gradients ride on the dgrad GEMM as one extra per-sequence k-block
                main.wait_event(ev_a)
                ops.gemm(dqkv, W["Wqkv"], out=dx, b_mn_major=True, residual=d_pre1, xk=(*sc.xk, Lp))
            else:
                ops.gemm(dqkv, W["Wqkv"], out=dx, b_mn_major=True, residual=d_pre1)
            if self._debug_skip_global or fused_dx:
                pass
            elif self.overlap_global:
                main.wait_event(ev_a)
                ops.global_attn_bwd_dx(*gargs, sv.glob[i], dx, sc.gws, drop_p=sv.drop_attn,
                                       drop_seed=self._seed(sv, i, 2))
            else:
                ops.global_attn_bwd(*gargs, sc.dctx, sv.glob[i], dx, G["Wqg"], G["bqg"], G["Wkg"], G["Wvg"], G["bvg"],
                                    ws=sc.gws, drop_p=sv.drop_attn, drop_seed=self._seed(sv, i, 2))
            d_out = dx
            aux_marks[i] = self.aux_done
            if ev_w is not None:               # this layer's *_global weight gradients (long finished: launched a layer ago)
                main.wait_event(ev_w)
            if self.grad_hook is not None:
                self.grad_hook(i)
        if self.aux_done is not None:         # join the aux stream: every gradient is final when backward() returns
            main.wait_event(self.aux_done)
        e = "embeddings."
        named = P._named
        gview = lambda k: P.view(k, P.grad) if named[k].requires_grad else None
        input_ids, token_type_ids, item_position_ids = sv.inputs
        ops.embed_ln_bwd(d_out, input_ids, token_type_ids, item_position_ids, sv.pos_ids,
                         P.view(e + "word_embeddings.weight"), P.view(e + "position_embeddings.weight"),
                         P.view(e + "token_type_embeddings.weight"), P.view(e + "item_position_embeddings.weight"),
                         P.view(e + "LayerNorm.weight"), P.view(e + "LayerNorm.bias"), Lp, cfg.pad_token_id,
                         cfg.layer_norm_eps, gview(e + "word_embeddings.weight"),
                         gview(e + "position_embeddings.weight"), gview(e + "token_type_embeddings.weight"),
                         gview(e + "item_position_embeddings.weight"), gview(e + "LayerNorm.weight"),
                         gview(e + "LayerNorm.bias"), drop_p=sv.drop_hidden, drop_seed=self._seed(sv, -1, 0))
