"""The reference's own training / evaluation loops driving the drop-in.

`finetune.py` imports `from recformer import ...` and then runs `train_one_epoch` (ref: finetune.py:98-137:
`autocast()`, `GradScaler.scale(loss).backward()`, `scaler.step(torch.optim.AdamW)`, `optimizer.zero_grad()`,
gradient accumulation, LambdaLR warm-up from ref: optimization.py:7-34) and `eval` (ref: finetune.py:66-96:
`scores = model(**batch)` -> `Ranker`).  The loops below restate those call sequences step by step on the `recformer`
alias package and compare the loss trajectory / metrics with the CPU oracle driven by the same torch optimiser."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import recformer_oracle as O

DEV = "cuda"
NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")      # ref: optimization.py:25


def _optimizer_and_scheduler(named_params, lr, weight_decay, warmup, total):
    """ref: optimization.py:22-34 (AdamW groups + linear warm-up / decay)."""
    named_params = list(named_params)
    groups = [{"params": [p for n, p in named_params if not any(nd in n for nd in NO_DECAY)], "weight_decay": weight_decay},
              {"params": [p for n, p in named_params if any(nd in n for nd in NO_DECAY)], "weight_decay": 0.0}]
    opt = torch.optim.AdamW(groups, lr=lr)
    lam = lambda s: float(s) / float(max(1, warmup)) if s < warmup else max(0.0, 1 - float(s) / float(max(1, total)))
    return opt, torch.optim.lr_scheduler.LambdaLR(opt, lam)


def _setup():
    from recformer import RecformerConfig, RecformerForSeqRec          # the reference's import line (finetune.py:12)
    cfg_kw = dict(vocab_size=1500, num_hidden_layers=2, attention_window=[64, 64], max_position_embeddings=600)
    ocfg = O.OracleConfig(**cfg_kw)
    config = RecformerConfig.from_pretrained("allenai/longformer-base-4096", **{k: v for k, v in cfg_kw.items()
                                                                                if k != "attention_window"},
                                             hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    config.max_attr_num, config.max_attr_length = 3, 32                 # ref: finetune.py:203-209
    config.max_item_embeddings, config.attention_window, config.max_token_num = 51, [64] * 2, 1024
    config.item_num, config.finetune_negative_sample_size = 60, 0
    model = RecformerForSeqRec(config)
    sd = O.make_state_dict(ocfg, seed=9, prefix="longformer.")
    model.load_state_dict(sd, strict=False)                             # ref: finetune.py:268
    args = types.SimpleNamespace(device=torch.device(DEV), fp16=True, gradient_accumulation_steps=2, metric_ks=[10, 50],
                                 learning_rate=5e-5, weight_decay=0.01, warmup_steps=1)
    model.to(args.device)
    items = O.make_item_table(config.item_num, 768, seed=1)
    model.init_item_embedding(items.clone())                            # ref: finetune.py:298 (CPU tensor) ...
    model.to(args.device)                                               # ... then :300 "send item embeddings to device"
    return model, ocfg, sd, items, args


def test_reference_train_loop_on_dropin_matches_oracle_trajectory():
    model, ocfg, sd, items, args = _setup()
    n_micro = 8
    loader = []
    for s in range(n_micro):
        b = O.make_batch(ocfg, 3, 200, seed=40 + s % 2, ragged=True)          # two micro-batches, cycled
        b["labels"] = torch.tensor([(7 * (s % 2) + 3 * r) % 60 for r in range(3)])
        loader.append(b)
    total = n_micro // args.gradient_accumulation_steps
    # ---- oracle: same optimiser / schedule / accumulation on the CPU restatement ----
    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    oopt, osched = _optimizer_and_scheduler([(k, v) for k, v in osd.items() if v.is_floating_point()], args.learning_rate,
                                            args.weight_decay, args.warmup_steps, total)
    ref_losses = []
    for step, batch in enumerate(loader):
        feats = {k: v for k, v in batch.items() if k != "labels"}
        loss = O.seqrec_forward(osd, ocfg, feats, items, labels=batch["labels"])
        ref_losses.append(loss.item())
        (loss / args.gradient_accumulation_steps).backward()
        if (step + 1) % args.gradient_accumulation_steps == 0:
            oopt.step()               # the fp16 branch's order (finetune.py:121-128): optimiser, zero_grad, scheduler
            oopt.zero_grad()
            osched.step()
    # ---- the reference's train_one_epoch body (fp16 branch) on the drop-in ----
    optimizer, scheduler = _optimizer_and_scheduler(model.named_parameters(), args.learning_rate, args.weight_decay,
                                                    args.warmup_steps, total)
    scaler = torch.amp.GradScaler("cuda")
    model.train()
    losses = []
    for step, batch in enumerate(loader):
        batch = {k: v.to(args.device) for k, v in batch.items()}
        with torch.autocast("cuda"):
            loss = model(**batch)
        losses.append(loss.item())
        if args.gradient_accumulation_steps > 1:
            loss = loss / args.gradient_accumulation_steps
        scaler.scale(loss).backward()
        if (step + 1) % args.gradient_accumulation_steps == 0:
            scale_before = scaler.get_scale()
            scaler.step(optimizer)
            scaler.update()
            optimizer_was_run = scale_before <= scaler.get_scale()
            optimizer.zero_grad()
            assert optimizer_was_run            # bf16 operands / fp32 accumulation: the 65536x loss scale never overflows
            scheduler.step()
    print("drop-in losses", [round(x, 4) for x in losses], "oracle", [round(x, 4) for x in ref_losses])
    # micro-batches 0-3 run on the initial weights (the warm-up's first optimiser step has lr 0): logits-level tolerance;
    # 4-7 follow one / two real AdamW steps (lr 3.75e-5, 2.5e-5; the oracle's loss falls 4.56 -> 3.77 -> 3.30), where
    # weights whose gradient is summation noise move by +-lr in either implementation: looser, but a wrong gradient
    # scale / unscale or a stale bf16 weight shadow after torch.optim.AdamW's in-place update would miss the descent
    for a, r in zip(losses[:4], ref_losses[:4]):
        assert abs(a - r) < 2e-2, (losses, ref_losses)
    for a, r in zip(losses[4:], ref_losses[4:]):
        assert abs(a - r) < 0.1, (losses, ref_losses)
    assert ref_losses[0] - ref_losses[4] > 0.5 and ref_losses[4] - ref_losses[6] > 0.2


def test_reference_eval_loop_on_dropin_matches_oracle_metrics():
    from recformer_b200 import Ranker                                   # ref: utils.py:76-107 (same class name / call)
    model, ocfg, sd, items, args = _setup()
    model.eval()
    ranker = Ranker(args.metric_ks)
    sums, ref_sums, n_batches = None, None, 0
    for s in range(3):
        batch = O.make_batch(ocfg, 4, 150 + 20 * s, seed=70 + s, ragged=True)
        with torch.no_grad():
            ref_scores = O.seqrec_forward(sd, ocfg, batch, items)
        # labels at a rank whose score is separated from both neighbours by more than twice the logit tolerance
        top = torch.topk(ref_scores, 31, dim=-1)
        gaps = torch.minimum(top.values[:, :-2] - top.values[:, 1:-1], top.values[:, 1:-1] - top.values[:, 2:])
        pick = gaps.argmax(-1) + 1
        assert (gaps.max(-1).values > 4e-2).all()
        labels = top.indices[torch.arange(4), pick].unsqueeze(-1)                    # (B, 1) as the eval collator emits
        dev_batch = {k: v.to(args.device) for k, v in batch.items()}
        with torch.no_grad():
            scores = model(**dev_batch)                                                # ref: finetune.py:80
        res = ranker(scores, labels.to(args.device))                                  # ref: finetune.py:82
        ref = O.ranker(ref_scores, labels, ks=tuple(args.metric_ks))
        assert (scores.cpu() - ref_scores).abs().max() < 2e-2
        sums = res if sums is None else [a + b for a, b in zip(sums, res)]
        ref_sums = ref if ref_sums is None else [a + b for a, b in zip(ref_sums, ref)]
        n_batches += 1
    got = [v / n_batches for v in sums]
    want = [v / n_batches for v in ref_sums]
    for i in range(4):       # NDCG@10, Recall@10, NDCG@50, Recall@50: equal to 4 decimals
        assert round(got[i], 4) == round(want[i], 4), (got, want)
    # a final batch of ONE user (len(dataset) % batch_size == 1) must not crash the metric (the reference's CE
    # raises on the squeezed 0-d target and is caught upstream, utils.py:84-89)
    one = {k: v[:1].to(args.device) for k, v in batch.items()}
    with torch.no_grad():
        res1 = ranker(model(**one), labels[:1].to(args.device))
    assert len(res1) == 2 * len(args.metric_ks) + 3 and all(v == v for v in res1)
