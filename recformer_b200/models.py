"""Drop-in replacements for the reference's hot-path classes (ref: recformer/models.py):
RecformerModel (:174-356), RecformerForSeqRec (:524-599), Similarity (:358-369), with the same
constructor, forward kwargs, outputs and `state_dict` keys, executed by the
B200 kernels behind the C ABI (include/recformer_b200.h).  The module tree below only CARRIES
parameters under the reference's names; all arithmetic is in recformer_b200/engine.py + csrc/.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .config import RecformerConfig
from .engine import EncoderEngine


# --------------------------------------------------------------------------------------------
# output container (ref returns transformers' LongformerBaseModelOutputWithPooling)
# --------------------------------------------------------------------------------------------
@dataclass
class RecformerModelOutput:
    last_hidden_state: torch.Tensor = None
    pooler_output: torch.Tensor = None
    hidden_states: Optional[Tuple[torch.Tensor]] = None
    attentions: Optional[Tuple[torch.Tensor]] = None
    global_attentions: Optional[Tuple[torch.Tensor]] = None

    def to_tuple(self):
        return tuple(v for v in (self.last_hidden_state, self.pooler_output, self.hidden_states, self.attentions,
                                 self.global_attentions) if v is not None)

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return self.to_tuple()[k]


# --------------------------------------------------------------------------------------------
# parameter containers mirroring the reference / HF module names
# --------------------------------------------------------------------------------------------
class RecformerEmbeddings(nn.Module):
    """ref: recformer/models.py:82-106 (forward is fused into rf_embed_ln_fwd)."""

    def __init__(self, config: RecformerConfig):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size,
                                                padding_idx=config.pad_token_id)
        self.token_type_embeddings = nn.Embedding(config.token_type_size, config.hidden_size)
        self.item_position_embeddings = nn.Embedding(config.max_item_embeddings, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.padding_idx = config.pad_token_id


class _SelfAttention(nn.Module):   # HF:445-466
    def __init__(self, config):
        super().__init__()
        E = config.hidden_size
        self.query, self.key, self.value = nn.Linear(E, E), nn.Linear(E, E), nn.Linear(E, E)
        self.query_global, self.key_global, self.value_global = nn.Linear(E, E), nn.Linear(E, E), nn.Linear(E, E)


class _SelfOutput(nn.Module):      # HF:1060-1066
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class _Attention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = _SelfAttention(config)
        self.output = _SelfOutput(config)


class _Intermediate(nn.Module):    # HF:1103-1111
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class _Output(nn.Module):          # HF:1119-1125
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class _Layer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = _Attention(config)
        self.intermediate = _Intermediate(config)
        self.output = _Output(config)


class _Encoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(config) for _ in range(config.num_hidden_layers)])


class RecformerPooler(nn.Module):
    """ref: recformer/models.py:155-171."""

    def __init__(self, config: RecformerConfig):
        super().__init__()
        self.pooler_type = config.pooler_type

    def forward(self, attention_mask: torch.Tensor, hidden_states: torch.Tensor) -> torch.Tensor:
        if self.pooler_type == "cls":
            return hidden_states[:, 0]
        if self.pooler_type == "avg":
            # the reference multiplies by the MERGED mask (0 padding / 1 local / 2 global, ref: recformer/models.py:310-312,
            # 342), so the CLS row counts twice in numerator and denominator.  Rows with mask 0 are zeroed by selection,
            # not by the product: with padding_rows_unused the encoder leaves them undefined (possibly non-finite).
            am = attention_mask[:, : hidden_states.shape[1]]
            h = hidden_states.masked_fill((am == 0).unsqueeze(-1), 0.0)
            am = am.to(hidden_states.dtype)
            return (h * am.unsqueeze(-1)).sum(1) / am.sum(-1).unsqueeze(-1)
        raise NotImplementedError


def _init_weights(module: nn.Module, std: float):
    """HF PreTrainedModel._init_weights for the module kinds used here."""
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=std)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=std)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)


# --------------------------------------------------------------------------------------------
# autograd bridge: one Function for the whole encoder
# --------------------------------------------------------------------------------------------
class _EncoderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hook, model, L, inputs):
        eng = model._engine
        sv = eng.forward(*inputs, training=model.training, save=True, skip_padding=model.padding_rows_unused)
        B, Lp, E = sv.B, sv.Lp, model.config.hidden_size
        hidden = eng.hidden(sv).view(B, Lp, E)[:, :L].clone()   # fp32 residual stream of the last layer
        ctx.model, ctx.sv, ctx.L = model, sv, L
        return hidden

    @staticmethod
    def backward(ctx, d_hidden):
        model, sv = ctx.model, ctx.sv
        eng = model._engine
        B, Lp, E = sv.B, sv.Lp, model.config.hidden_size
        dout = torch.zeros(B, Lp, E, dtype=torch.bfloat16, device=d_hidden.device)
        dout[:, : ctx.L] = d_hidden
        eng.backward(sv, dout.view(B * Lp, E))
        eng.release(sv)
        ctx.sv = None
        return None, None, None, None


class _EncoderPooledFunction(torch.autograd.Function):
    """The encoder when only the 'cls' pooler output is consumed (RecformerForSeqRec / the contrastive towers): returns
    the [B, E] CLS rows, so neither the 50 MB fp32 copy of the hidden states nor a per-step zero-filled upstream
    gradient is materialised."""

    @staticmethod
    def forward(ctx, hook, model, inputs):
        eng = model._engine
        sv = eng.forward(*inputs, training=model.training, save=True, skip_padding=True)   # only CLS rows leave / return
        pooled = eng.hidden(sv).view(sv.B, sv.Lp, model.config.hidden_size)[:, 0].clone()
        ctx.model, ctx.sv = model, sv
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        eng = ctx.model._engine
        eng.backward_from_pooled(ctx.sv, d_pooled)
        eng.release(ctx.sv)
        ctx.sv = None
        return None, None, None


class RecformerModel(nn.Module):
    """ref: recformer/models.py:174-356 — same constructor checks, forward kwargs and outputs."""

    def __init__(self, config: RecformerConfig):
        super().__init__()
        self.config = config
        if isinstance(config.attention_window, int):
            assert config.attention_window % 2 == 0, "`config.attention_window` has to be an even value"
            assert config.attention_window > 0, "`config.attention_window` has to be positive"
            config.attention_window = [config.attention_window] * config.num_hidden_layers
        else:
            assert len(config.attention_window) == config.num_hidden_layers, (
                "`len(config.attention_window)` should equal `config.num_hidden_layers`. "
                f"Expected {config.num_hidden_layers}, given {len(config.attention_window)}")
        if config.hidden_size != 768 or config.num_attention_heads != 12:
            raise NotImplementedError("recformer_b200 kernels are built for the longformer-base shape "
                                      "(hidden 768, 12 heads of 64)")
        self.embeddings = RecformerEmbeddings(config)
        self.encoder = _Encoder(config)
        self.pooler = RecformerPooler(config)
        self.apply(lambda m: _init_weights(m, config.initializer_range))
        # not parameters / buffers: the engine and the autograd hook stay out of state_dict
        object.__setattr__(self, "_engine", EncoderEngine(self))
        object.__setattr__(self, "_hook", None)
        self.strict_checks = True
        # True: whoever calls forward() never reads `last_hidden_state` at padded positions and never sends gradient to
        # them (the pretraining heads: CLS rows + gathered masked rows) — the encoder may then skip 256-row tiles made of
        # padding only (engine.forward(skip_padding=True)); those rows of the returned hidden states are undefined.
        # False (default): every position is computed, as in the reference (ref: recformer/models.py:274-356).
        self.padding_rows_unused = False

    # HF-compatible accessors used by the reference scripts (finetune.py:272-275)
    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        # the engine re-adopts the new Parameter on the next forward (FlatParams.ensure compares identities)
        self.embeddings.word_embeddings = value

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._engine.params.invalidate()        # copy_() under no_grad may not bump the version counters
        return out

    def named_parameters(self, *a, **kw):   # engine views must see the module's own parameters
        return super().named_parameters(*a, **kw)

    def _grad_hook(self, device):
        if self._hook is None or self._hook.device != torch.device(device):
            object.__setattr__(self, "_hook", torch.zeros(1, device=device, requires_grad=True))
        return self._hook

    def forward(self,
                input_ids: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None,
                global_attention_mask: Optional[torch.Tensor] = None,
                head_mask: Optional[torch.Tensor] = None,
                token_type_ids: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None,
                item_position_ids: Optional[torch.Tensor] = None,
                inputs_embeds: Optional[torch.Tensor] = None,
                output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None):
        return_dict = return_dict if return_dict is not None else self.config.use_return_dict
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        if inputs_embeds is not None:
            raise NotImplementedError("recformer_b200: inputs_embeds is not on the accelerated path")
        if head_mask is not None:
            raise NotImplementedError("recformer_b200: head_mask is not supported (the reference never sets it)")
        if output_attentions or output_hidden_states:
            raise NotImplementedError("recformer_b200: attention maps / per-layer states are never materialised")
        if not input_ids.is_cuda:
            raise RuntimeError("recformer_b200 runs on CUDA only (no CPU fallback): move inputs to the GPU")
        B, L = input_ids.shape
        if item_position_ids is None:
            raise ValueError("item_position_ids is required (ref: recformer/models.py:132 indexes it unconditionally)")
        to64 = lambda t: None if t is None else t.to(torch.int64).contiguous()
        inputs = (to64(input_ids), to64(attention_mask), to64(global_attention_mask), to64(token_type_ids),
                  to64(item_position_ids), to64(position_ids))
        eng = self._engine
        E = self.config.hidden_size
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            hidden = _EncoderFunction.apply(self._grad_hook(input_ids.device), self, L, inputs)
        else:
            sv = eng.forward(*inputs, training=self.training, save=False, skip_padding=self.padding_rows_unused)
            hidden = eng.hidden(sv).view(B, sv.Lp, E)[:, :L].clone()
            eng.release(sv)
        if self.strict_checks:
            eng.check_errors()
        merged = None
        if self.pooler.pooler_type != "cls":
            am = attention_mask if attention_mask is not None else torch.ones_like(input_ids)
            merged = am * (global_attention_mask + 1) if global_attention_mask is not None else am
        pooled = self.pooler(merged, hidden)
        if not return_dict:
            return (hidden, pooled)
        return RecformerModelOutput(last_hidden_state=hidden, pooler_output=pooled)

    def forward_pooled(self, input_ids, attention_mask=None, global_attention_mask=None, token_type_ids=None,
                       position_ids=None, item_position_ids=None) -> torch.Tensor:
        """`forward(...).pooler_output` for the 'cls' pooler without materialising `last_hidden_state` (same checks,
        same values; what RecformerForSeqRec and the pretraining towers consume)."""
        if self.pooler.pooler_type != "cls":
            return self(input_ids, attention_mask=attention_mask, global_attention_mask=global_attention_mask,
                        token_type_ids=token_type_ids, position_ids=position_ids, item_position_ids=item_position_ids,
                        return_dict=True).pooler_output
        if not input_ids.is_cuda:
            raise RuntimeError("recformer_b200 runs on CUDA only (no CPU fallback): move inputs to the GPU")
        if item_position_ids is None:
            raise ValueError("item_position_ids is required (ref: recformer/models.py:132 indexes it unconditionally)")
        to64 = lambda t: None if t is None else t.to(torch.int64).contiguous()
        inputs = (to64(input_ids), to64(attention_mask), to64(global_attention_mask), to64(token_type_ids),
                  to64(item_position_ids), to64(position_ids))
        eng = self._engine
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            pooled = _EncoderPooledFunction.apply(self._grad_hook(input_ids.device), self, inputs)
        else:
            sv = eng.forward(*inputs, training=self.training, save=False, skip_padding=True)
            pooled = eng.hidden(sv).view(sv.B, sv.Lp, self.config.hidden_size)[:, 0].clone()
            eng.release(sv)
        if self.strict_checks:
            eng.check_errors()
        return pooled

    # raw bf16 access for fused consumers (scoring) that do not need fp32 copies
    def encode_pooled_bf16(self, **batch) -> torch.Tensor:
        """CLS vectors as bf16 [B,E] without materialising fp32 hidden states (inference only)."""
        eng = self._engine
        to64 = lambda t: None if t is None else t.to(torch.int64).contiguous()
        ids = batch["input_ids"]
        sv = eng.forward(to64(ids), to64(batch.get("attention_mask")), to64(batch.get("global_attention_mask")),
                         to64(batch.get("token_type_ids")), to64(batch["item_position_ids"]),
                         to64(batch.get("position_ids")), training=False, save=False, skip_padding=True)
        pooled = eng.hidden_bf16(sv).view(sv.B, sv.Lp, -1)[:, 0].contiguous()
        eng.release(sv)
        return pooled


# --------------------------------------------------------------------------------------------
# scoring
# --------------------------------------------------------------------------------------------
class _CosineCEFunction(torch.autograd.Function):
    """loss = CE(cos(pooled, items)/temp, labels) with the table pre-normalised (Spec S)."""

    @staticmethod
    def forward(ctx, pooled, yn, labels, temp):
        loss, dpooled = ops.cosine_ce(pooled.contiguous(), yn, labels, temp, want_grad=True)
        ctx.save_for_backward(dpooled)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        (dpooled,) = ctx.saved_tensors
        return dpooled * g, None, None, None


class _CandidateCEFunction(torch.autograd.Function):
    """Sampled-softmax CE over (label, negatives) candidates (ref: recformer/models.py:593-597)."""

    @staticmethod
    def forward(ctx, pooled, yn, cand, temp):
        loss, dpooled = ops.cosine_candidates_ce(pooled.detach().float().contiguous(), yn, cand.contiguous(), temp)
        ctx.save_for_backward(dpooled)
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, g):
        (dpooled,) = ctx.saved_tensors
        return dpooled * g, None, None, None


class Similarity(nn.Module):
    """ref: recformer/models.py:358-369 — cos(x, y) / temp on broadcastable (B,1,E) x (1|B,N,E)."""

    def __init__(self, config: RecformerConfig):
        super().__init__()
        self.temp = config.temp

    def forward(self, x, y):
        if x.dim() == 3 and x.shape[1] == 1 and y.dim() == 3 and y.shape[0] == 1:
            xn = ops.normalize_rows(x[:, 0].contiguous())
            yn = ops.normalize_rows(y[0].contiguous())
            return ops.cosine_logits(xn, yn, self.temp)
        xn = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-8)
        yn = y / y.norm(dim=-1, keepdim=True).clamp_min(1e-8)
        return (xn * yn).sum(-1) / self.temp


class RecformerForSeqRec(nn.Module):
    """ref: recformer/models.py:524-599."""

    def __init__(self, config: RecformerConfig):
        super().__init__()
        self.config = config
        self.longformer = RecformerModel(config)
        self.sim = Similarity(config)
        object.__setattr__(self, "_yn", None)      # cached L2-normalised bf16 table
        object.__setattr__(self, "_yn_sig", None)

    def init_item_embedding(self, embeddings: Optional[torch.Tensor] = None):
        self.item_embedding = nn.Embedding(num_embeddings=self.config.item_num, embedding_dim=self.config.hidden_size)
        if embeddings is not None:
            self.item_embedding = nn.Embedding.from_pretrained(embeddings, freeze=True)
            print("Initalize item embeddings from vectors.")
        object.__setattr__(self, "_yn_sig", None)

    def normalized_items(self) -> torch.Tensor:
        """bf16 L2-normalised copy of the (frozen) item table, rebuilt only when the table changes."""
        w = self.item_embedding.weight
        sig = (w.data_ptr(), w._version, tuple(w.shape), str(w.device))
        if self._yn is None or self._yn_sig != sig:
            object.__setattr__(self, "_yn", ops.normalize_rows(w.detach().contiguous()))
            object.__setattr__(self, "_yn_sig", sig)
        return self._yn

    def similarity_score(self, pooler_output, candidates=None):
        if candidates is None:
            xn = ops.normalize_rows(pooler_output.detach().contiguous())
            return ops.cosine_logits(xn, self.normalized_items(), self.config.temp)
        if candidates.dim() == 2 and pooler_output.is_cuda:          # gather + dot fused: no (B, C, E) tensor
            return ops.cosine_candidates(pooler_output.detach().float().contiguous(), self.normalized_items(),
                                         candidates.to(torch.int64).contiguous(), self.config.temp)
        candidate_embeddings = self.item_embedding(candidates)
        return self.sim(pooler_output.unsqueeze(1), candidate_embeddings)

    @torch.no_grad()
    def topk(self, pooler_output, k: int = 10, labels: Optional[torch.Tensor] = None, id_base: int = 0):
        """Fused full-catalogue scoring + top-k (the (B,N) logits are never written)."""
        xn = ops.normalize_rows(pooler_output.contiguous())
        return ops.cosine_topk(xn, self.normalized_items(), self.config.temp, k=k, id_base=id_base, labels=labels)

    def forward(self,
                input_ids: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None,
                global_attention_mask: Optional[torch.Tensor] = None,
                head_mask: Optional[torch.Tensor] = None,
                token_type_ids: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None,
                item_position_ids: Optional[torch.Tensor] = None,
                inputs_embeds: Optional[torch.Tensor] = None,
                output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None,
                candidates: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None):
        batch_size = input_ids.size(0)
        if head_mask is None and inputs_embeds is None and not output_attentions and not output_hidden_states:
            # only the CLS rows are consumed (ref :565-579 reads outputs.pooler_output)
            pooler_output = self.longformer.forward_pooled(input_ids, attention_mask=attention_mask,
                                                           global_attention_mask=global_attention_mask,
                                                           token_type_ids=token_type_ids, position_ids=position_ids,
                                                           item_position_ids=item_position_ids)
        else:                       # unsupported kwargs raise inside RecformerModel.forward, as documented there
            pooler_output = self.longformer(input_ids, attention_mask=attention_mask,
                                            global_attention_mask=global_attention_mask, head_mask=head_mask,
                                            token_type_ids=token_type_ids, position_ids=position_ids,
                                            item_position_ids=item_position_ids, inputs_embeds=inputs_embeds,
                                            output_attentions=output_attentions,
                                            output_hidden_states=output_hidden_states, return_dict=True).pooler_output
        if labels is None:
            return self.similarity_score(pooler_output, candidates)
        if self.config.finetune_negative_sample_size <= 0:      # full softmax
            return _CosineCEFunction.apply(pooler_output, self.normalized_items(), labels.to(torch.int64).contiguous(),
                                           self.config.temp)
        # sampled softmax (ref: :593-597) — negatives drawn on the device instead of CPU + H2D
        neg = torch.randint(0, self.config.item_num, (batch_size, self.config.finetune_negative_sample_size),
                            device=labels.device)
        cand = torch.cat((labels.unsqueeze(-1), neg), dim=-1).to(torch.int64)
        return _CandidateCEFunction.apply(pooler_output, self.normalized_items(), cand, self.config.temp)


# --------------------------------------------------------------------------------------------
# binary classification head on the CLS vector (ref: recformer/models.py:601-713; the caller finetune_classification.py)
# --------------------------------------------------------------------------------------------
class FocalLoss(nn.Module):
    """ref: recformer/models.py:601-631 — mean of alpha_t * (1 - p_t)^gamma * BCE-with-logits(inputs, targets)."""

    def __init__(self, alpha=1, gamma=2, pos_weight=None):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma
        self.pos_weight = pos_weight

    def forward(self, inputs, targets):
        probs = torch.sigmoid(inputs)
        ce = nn.functional.binary_cross_entropy_with_logits(inputs, targets, pos_weight=self.pos_weight, reduction="none")
        p_t = probs * targets + (1 - probs) * (1 - targets)
        loss = (1 - p_t) ** self.gamma * ce
        if self.alpha is not None:
            loss = (self.alpha * targets + (1 - self.alpha) * (1 - targets)) * loss
        return loss.mean()


class RecformerForFraudDetection(nn.Module):
    """ref: recformer/models.py:633-713 — dropout + a 768 -> 384 -> 192 -> 1 ReLU MLP on the encoder's pooled output,
    BCE-with-logits (config.pos_weight) against 0/1 labels.  The encoder pass is the CUDA engine (CLS-only entry when
    the pooler is 'cls': forward_pooled); the head works on (B, 768) vectors — a few MFLOP — and stays plain torch, its
    parameters live outside the encoder's flat buffer (FusedAdamW steps them through its `extra_params` optimizer)."""

    def __init__(self, config: RecformerConfig):
        super().__init__()
        self.config = config
        self.longformer = RecformerModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob if hasattr(config, "hidden_dropout_prob") else 0.1)
        E = config.hidden_size
        self.classifier = nn.Sequential(nn.Linear(E, E // 2), nn.ReLU(), nn.Dropout(0.2),
                                        nn.Linear(E // 2, E // 4), nn.ReLU(), nn.Dropout(0.2),
                                        nn.Linear(E // 4, 1))
        self.classifier.apply(lambda m: _init_weights(m, config.initializer_range))       # HF post_init (ref :659)
        # a plain attribute, not a buffer: absent from the state dict, as in the reference (ref :656)
        self.pos_weight = torch.tensor(config.pos_weight) if hasattr(config, "pos_weight") else torch.tensor(1.0)

    def init_item_embedding(self, embeddings: Optional[torch.Tensor] = None):
        self.item_embedding = nn.Embedding(num_embeddings=self.config.item_num, embedding_dim=self.config.hidden_size)
        if embeddings is not None:
            self.item_embedding = nn.Embedding.from_pretrained(embeddings, freeze=True)
            print("Initalize item embeddings from vectors.")

    def forward(self,
                input_ids: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None,
                global_attention_mask: Optional[torch.Tensor] = None,
                head_mask: Optional[torch.Tensor] = None,
                token_type_ids: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None,
                item_position_ids: Optional[torch.Tensor] = None,
                inputs_embeds: Optional[torch.Tensor] = None,
                output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None,
                labels: Optional[torch.Tensor] = None):
        return_dict = return_dict if return_dict is not None else self.config.use_return_dict
        if head_mask is None and inputs_embeds is None and not output_attentions and not output_hidden_states:
            pooler_output = self.longformer.forward_pooled(input_ids, attention_mask=attention_mask,
                                                           global_attention_mask=global_attention_mask,
                                                           token_type_ids=token_type_ids, position_ids=position_ids,
                                                           item_position_ids=item_position_ids)
        else:                       # unsupported kwargs raise inside RecformerModel.forward, as documented there
            pooler_output = self.longformer(input_ids, attention_mask=attention_mask,
                                            global_attention_mask=global_attention_mask, head_mask=head_mask,
                                            token_type_ids=token_type_ids, position_ids=position_ids,
                                            item_position_ids=item_position_ids, inputs_embeds=inputs_embeds,
                                            output_attentions=output_attentions,
                                            output_hidden_states=output_hidden_states, return_dict=True).pooler_output
        logits = self.classifier(self.dropout(pooler_output)).squeeze(-1)            # (B,)
        loss = None
        if labels is not None:
            loss = nn.functional.binary_cross_entropy_with_logits(logits, labels.float(),
                                                                  pos_weight=self.pos_weight.to(logits.device))
        if not return_dict:
            return (loss, logits)
        return {"loss": loss, "logits": logits}


# --------------------------------------------------------------------------------------------
# pretraining (ref: recformer/models.py:372-520)
# --------------------------------------------------------------------------------------------
@dataclass
class RecformerPretrainingOutput:
    """ref: recformer/models.py:57-65."""
    cl_correct_num: float = 0.0
    cl_total_num: float = 1e-5
    loss: Optional[torch.Tensor] = None
    logits: torch.Tensor = None
    hidden_states: Optional[Tuple[torch.Tensor]] = None
    attentions: Optional[Tuple[torch.Tensor]] = None
    global_attentions: Optional[Tuple[torch.Tensor]] = None


class _LMHead(nn.Module):
    """Parameter container of HF LongformerLMHead (HF:1264-1283): dense, layer_norm, decoder (+ the unused
    `bias` parameter that transformers 5.5.0 keeps in the state dict)."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.layer_norm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))


class _LMHeadCEFunction(torch.autograd.Function):
    """mean CE(decoder(LN(gelu(dense(rows)))), labels) over the MASKED rows only (the reference scores every
    position — (B,L,50265) fp32 — and lets ignore_index drop 85 % of it; same loss, ~6.7x less work).
    All GEMMs run on the tcgen05 kernel; the vocabulary is padded to a multiple of 32 columns."""

    @staticmethod
    def forward(ctx, rows, labels, Wd, bd, lnw, lnb, Wdec, bdec, eps):
        M, E = rows.shape
        V = Wdec.shape[0]
        Vp = (V + 31) // 32 * 32
        dev = rows.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        x = rows.detach().to(torch.bfloat16).contiguous()
        Wd16 = Wd.detach().to(torch.bfloat16).contiguous()
        u, g = torch.empty(M, E, **bf), torch.empty(M, E, **bf)
        ops.gemm(x, Wd16, out=u, bias=bd.detach().contiguous(), epi=ops.EPI_GELU, out2=g)   # u = gelu', g = gelu
        g32 = g.float()
        y = torch.empty(M, E, **bf)
        stats = torch.empty(M, 2, dtype=torch.float32, device=dev)
        ops.layernorm_fwd(g32, lnw.detach().contiguous(), lnb.detach().contiguous(), eps, out=y, stats=stats)
        Wdec16 = torch.zeros(Vp, E, **bf)
        Wdec16[:V] = Wdec.detach()
        bpad = torch.zeros(Vp, dtype=torch.float32, device=dev)
        bpad[:V] = bdec.detach()
        logits = torch.empty(M, Vp, dtype=torch.float32, device=dev)
        ops.gemm(y, Wdec16, out=logits, bias=bpad)
        loss, dlogits = ops.mlm_ce(logits, labels.contiguous(), V)
        del logits
        ctx.save_for_backward(x, Wd16, u, g32, stats, y, Wdec16, dlogits, lnw.detach())
        ctx.V = V
        return loss.squeeze(0)

    @staticmethod
    def backward(ctx, gout):
        x, Wd16, u, g32, stats, y, Wdec16, dlogits, lnw = ctx.saved_tensors
        M, E = x.shape
        V, Vp = ctx.V, Wdec16.shape[0]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        dlogits = dlogits * gout.to(torch.bfloat16)            # upstream scale (mlm_weight, grad accumulation, ...)
        dbpad = torch.zeros(Vp, **f32)
        ops.colsum(dlogits, dbpad)
        dWdec = torch.zeros(Vp, E, **f32)
        ops.gemm(dlogits, y, out=dWdec, a_mn_major=True, b_mn_major=True, accumulate=True)
        dy = torch.empty(M, E, dtype=torch.bfloat16, device=dev)
        ops.gemm(dlogits, Wdec16, out=dy, b_mn_major=True)
        dlnw, dlnb = torch.zeros(E, **f32), torch.zeros(E, **f32)
        dg = ops.layernorm_bwd(dy, g32, stats, lnw.contiguous(), dlnw, dlnb)
        dpre = (dg * u).contiguous()
        dbd = torch.zeros(E, **f32)
        ops.colsum(dpre, dbd)
        dWd = torch.zeros(E, E, **f32)
        ops.gemm(dpre, x, out=dWd, a_mn_major=True, b_mn_major=True, accumulate=True)
        dx = torch.empty(M, E, dtype=torch.bfloat16, device=dev)
        ops.gemm(dpre, Wd16, out=dx, b_mn_major=True)
        return dx.float(), None, dWd, dbd, dlnw, dlnb, dWdec[:V], dbpad[:V], None


def gather_cls_with_local_grad(z1: torch.Tensor, z2: torch.Tensor, group=None):
    """ref: recformer/models.py:475-490 — every rank's CLS vectors become in-batch negatives; the gathered copies
    carry no gradient, so this rank's slot is replaced by the live tensors.  ONE collective on the packed
    (2, B, E) buffer instead of the reference's two list all_gathers; device-agnostic (NCCL on GPUs, gloo in the
    CPU tests).  Returns (z1_all, z2_all) of shape (B * world, E), rows ordered by rank."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    packed = torch.stack([z1.detach(), z2.detach()]).contiguous()
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1), group=group)
    parts1 = [gathered[r, 0] if r != rank else z1 for r in range(world)]
    parts2 = [gathered[r, 1] if r != rank else z2 for r in range(world)]
    return torch.cat(parts1, 0), torch.cat(parts2, 0)


def contrastive_head(z1: torch.Tensor, z2: torch.Tensor, temp: float):
    """ref: recformer/models.py:492-497 — cos_sim = sim(z1[:,None], z2[None]) (Spec S), CE against the diagonal.
    (B*W)^2 logits on [B*W, 768] vectors: fp32 keeps both towers' gradients exact.  Returns (loss, cos_sim, correct)."""
    z1n = z1 / z1.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    z2n = z2 / z2.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    cos_sim = (z1n @ z2n.t()) / temp
    target = torch.arange(cos_sim.size(0), device=cos_sim.device)
    loss = nn.functional.cross_entropy(cos_sim, target)
    return loss, cos_sim, (torch.argmax(cos_sim, 1) == target).sum()


class RecformerForPretraining(nn.Module):
    """ref: recformer/models.py:372-520 — two-tower in-batch contrastive loss on the CLS vectors of (history a,
    target item b) plus mlm_weight * masked-LM loss on the masked copies of both; same forward kwargs, output
    fields and state_dict keys (`longformer.*`, `lm_head.*`)."""

    def __init__(self, config: RecformerConfig):
        super().__init__()
        self.config = config
        self.longformer = RecformerModel(config)
        self.longformer.padding_rows_unused = True      # this class consumes CLS rows and masked (real-token) rows only
        self.lm_head = _LMHead(config)
        self.lm_head.apply(lambda m: _init_weights(m, config.initializer_range))
        self.sim = Similarity(config)

    def _encode(self, ids, tag, kw):
        return self.longformer(ids, attention_mask=kw.get(f"attention_mask_{tag}"),
                               global_attention_mask=kw.get(f"global_attention_mask_{tag}"),
                               token_type_ids=kw.get(f"token_type_ids_{tag}"),
                               item_position_ids=kw.get(f"item_position_ids_{tag}"), return_dict=True)

    def _mlm_loss(self, hidden, labels):
        flat = labels.reshape(-1)
        idx = torch.nonzero(flat >= 0).squeeze(1)             # masked positions (one host sync per call)
        if idx.numel() == 0:
            return hidden.sum() * 0.0
        rows = hidden.reshape(-1, hidden.shape[-1]).index_select(0, idx)
        h = self.lm_head
        return _LMHeadCEFunction.apply(rows, flat.index_select(0, idx), h.dense.weight, h.dense.bias, h.layer_norm.weight,
                                       h.layer_norm.bias, h.decoder.weight, h.decoder.bias, self.config.layer_norm_eps)

    def forward(self, input_ids_a=None, attention_mask_a=None, global_attention_mask_a=None, token_type_ids_a=None,
                item_position_ids_a=None, mlm_input_ids_a=None, mlm_labels_a=None, input_ids_b=None,
                attention_mask_b=None, global_attention_mask_b=None, token_type_ids_b=None, item_position_ids_b=None,
                mlm_input_ids_b=None, mlm_labels_b=None, head_mask=None, position_ids=None, inputs_embeds=None,
                labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        import torch.distributed as dist
        if head_mask is not None or inputs_embeds is not None or position_ids is not None:
            raise NotImplementedError("recformer_b200: head_mask / inputs_embeds / position_ids are not supported here")
        kw = dict(attention_mask_a=attention_mask_a, global_attention_mask_a=global_attention_mask_a,
                  token_type_ids_a=token_type_ids_a, item_position_ids_a=item_position_ids_a,
                  attention_mask_b=attention_mask_b, global_attention_mask_b=global_attention_mask_b,
                  token_type_ids_b=token_type_ids_b, item_position_ids_b=item_position_ids_b)
        batch_size = input_ids_a.size(0)
        pooled = lambda ids, tag: self.longformer.forward_pooled(
            ids, attention_mask=kw.get(f"attention_mask_{tag}"), global_attention_mask=kw.get(f"global_attention_mask_{tag}"),
            token_type_ids=kw.get(f"token_type_ids_{tag}"), item_position_ids=kw.get(f"item_position_ids_{tag}"))
        z1 = pooled(input_ids_a, "a")
        z2 = pooled(input_ids_b, "b")
        if dist.is_available() and dist.is_initialized() and self.training and dist.get_world_size() > 1:
            z1, z2 = gather_cls_with_local_grad(z1, z2)            # ref :475-490
        loss, cos_sim, correct_num = contrastive_head(z1, z2, self.config.temp)
        if mlm_input_ids_a is not None and mlm_labels_a is not None:
            hidden = self._encode(mlm_input_ids_a, "a", kw).last_hidden_state
            loss = loss + self.config.mlm_weight * self._mlm_loss(hidden, mlm_labels_a)
        if mlm_input_ids_b is not None and mlm_labels_b is not None:
            hidden = self._encode(mlm_input_ids_b, "b", kw).last_hidden_state
            loss = loss + self.config.mlm_weight * self._mlm_loss(hidden, mlm_labels_b)
        return RecformerPretrainingOutput(loss=loss, logits=cos_sim, cl_correct_num=correct_num, cl_total_num=batch_size)
