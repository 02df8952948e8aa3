"""Secondary workloads of BASELINE.json measured with the same device-timing rules as bench.py:
  configs[2]  pretraining step (MLM + in-batch contrastive, batch 64/GPU, data-parallel)  -> pretrain_bench
  configs[4]  long-sequence stress (2 x 4096 tokens, attention_window 64 -> 512, fwd+bwd) -> longseq_bench
bench.py carries their results as extra keys of its JSON line; tools/bench_pretrain.py / bench_longseq.py print them
stand-alone.  Synthetic weights / batches come from the seeded generators in tools/synthetic.py (not from the oracle)."""
from __future__ import annotations

import torch

E, NL = 768, 12
DENSE_FLOP_PER_TOKEN_LAYER = 14155776


def _timed(fn, steps, world):
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def pretrain_bench(device, rank, world, steps=3, warmup=3, B=64, LA=1024, LB=128, sustained_tflops=None):
    import recformer_b200 as rb
    from recformer_b200 import dist as rdist
    from recformer_b200.optim import FusedAdamW
    from tools import synthetic as O
    ocfg = O.SynthConfig()
    cfg = rb.RecformerConfig(attention_window=[64] * NL, max_token_num=LA, max_item_embeddings=51, max_attr_num=3,
                             max_attr_length=32)
    model = rb.RecformerForPretraining(cfg)
    model.load_state_dict(O.make_pretrain_state_dict(ocfg, seed=0), strict=True)
    model = model.to(device).train()
    model.longformer.strict_checks = False
    opt = FusedAdamW(model, lr=5e-5, weight_decay=0.01)          # also steps lm_head.* (parameters outside the flat buffer)
    sync = rdist.GradSync(model, passes_per_step=4) if world > 1 else None
    batches = [{k: v.to(device) for k, v in O.make_pretrain_batch(ocfg, B, LA, LB, seed=100 * rank + i).items()}
               for i in range(2)]
    last = {}

    def step(i):
        out = model(**batches[i % 2])
        opt.zero_grad()
        out.loss.backward()
        if sync is not None:
            sync.finish()
        opt.step(grad_scale=1.0 / world)
        last["loss"] = out.loss

    for i in range(max(3, warmup)):
        step(i)
    ms = _timed(step, steps, world)
    tokens = B * (2 * LA + 2 * LB)          # padded tokens through the encoder per step and rank
    masked = sum(int((batches[0][k] >= 0).sum()) for k in ("mlm_labels_a", "mlm_labels_b"))
    enc_flops = 3 * (DENSE_FLOP_PER_TOKEN_LAYER + 4 * 66 * E) * NL * tokens
    head_flops = 3 * 2 * masked * (E * E + E * cfg.vocab_size)
    res = {"workload": "pretraining step (BASELINE configs[2]): MLM + in-batch contrastive, 4 encoder passes, LM head on the "
                       f"masked rows, batch {B}/GPU x (1024 + 128) tokens, dropout 0.1, fwd+bwd+AdamW, dp{world}",
           "n_gpus": world, "steps": steps, "ms_per_step": ms, "pairs_per_s": world * B / (ms / 1e3),
           "masked_rows_per_step": masked,
           "algorithmic_tflops_per_gpu": (enc_flops + head_flops) / (ms / 1e3) / 1e12,
           "last_loss": float(last["loss"].item()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if sustained_tflops:
        res["tensor_roof_frac"] = res["algorithmic_tflops_per_gpu"] / sustained_tflops
    if sync is not None:
        model.longformer._engine.grad_hook = None
    del model, opt, batches
    torch.cuda.empty_cache()
    return res


def longseq_bench(device, windows=(64, 128, 256, 512), steps=3, warmup=3, B=2, L=4096, sustained_tflops=None):
    import recformer_b200 as rb
    from tools import synthetic as O
    out = []
    for window in windows:
        cfg = rb.RecformerConfig(attention_window=[window] * NL, max_token_num=L, hidden_dropout_prob=0.0,
                                 attention_probs_dropout_prob=0.0)
        model = rb.RecformerModel(cfg).to(device).train()
        model.strict_checks = False
        batch = {k: v.to(device) for k, v in O.make_batch(O.SynthConfig(attention_window=[window] * NL), B, L, seed=1,
                                                          ragged=True).items()}

        def step(_):
            model(**batch).pooler_output.float().square().sum().backward()

        for i in range(max(3, warmup)):
            step(i)
        ms = _timed(step, steps, 1)
        # algorithmic FLOPs: dense 14 155 776 + band 4*(window+2)*768 per token-layer, x3 for fwd+bwd
        flops = 3 * (DENSE_FLOP_PER_TOKEN_LAYER + 4 * (window + 2) * E) * NL * B * L
        r = {"attention_window": window, "B": B, "L": L, "ms_per_step": ms, "seqs_per_s": B / (ms / 1e3),
             "tokens_per_s": B * L / (ms / 1e3), "algorithmic_tflops": flops / (ms / 1e3) / 1e12}
        if sustained_tflops:
            r["tensor_roof_frac"] = r["algorithmic_tflops"] / sustained_tflops
        out.append(r)
        del model, batch
        torch.cuda.empty_cache()
    return out


def encode_items_bench(device, n_items=8192, batch_size=512, sustained_tflops=None):
    """SURVEY.md §8f-1 — `encode_all_items` (ref: finetune.py:38-63; run once per stage-1 epoch, :304-307): one-item
    sequences of 20..96 tokens (padded to the batch max, then to the 64-token window), 12 layers, CLS pooled, rows also
    written as the L2-normalised bf16 shard.  Timed with the device-side batch assembly and with the reference's host
    tokenizer loop."""
    import time
    import numpy as np
    import recformer_b200 as rb
    from recformer_b200.items import encode_all_items
    from recformer_b200.tokenization import DeviceItemStore
    cfg = rb.RecformerConfig(attention_window=[64] * NL, max_token_num=1024, max_item_embeddings=51, max_attr_num=3,
                             max_attr_length=32)
    model = rb.RecformerModel(cfg).to(device).eval()
    model.strict_checks = False
    tok = rb.RecformerTokenizer(cfg)
    rng = np.random.default_rng(0)
    items = {}
    for i in range(n_items):
        n = int(rng.integers(20, 97))
        items[i] = [rng.integers(3, 50265, size=n).tolist(), rng.integers(1, 3, size=n).tolist()]
    norm = torch.empty(n_items, E, dtype=torch.bfloat16, device=device)
    store = DeviceItemStore(cfg, items, device=device)
    out = {}
    for name, kw in (("device_assembly", dict(item_store=store)), ("host_tokenizer", dict())):
        encode_all_items(model, tok, items, batch_size=batch_size, normalized_out=norm, id_range=(0, 2 * batch_size), **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        encode_all_items(model, tok, items, batch_size=batch_size, normalized_out=norm, **kw)
        torch.cuda.synchronize()
        out[name + "_items_per_s"] = n_items / (time.perf_counter() - t0)
    tokens = n_items * 128                      # every batch pads to 97..128 tokens -> 128 after window padding
    flops = (DENSE_FLOP_PER_TOKEN_LAYER + 4 * 66 * E) * NL * tokens
    out.update(workload=f"encode_all_items: {n_items} one-item sequences (20-96 tokens), batch {batch_size}, 12 layers, forward only",
               algorithmic_tflops=flops * out["device_assembly_items_per_s"] / n_items / 1e12)
    if sustained_tflops:
        out["tensor_roof_frac"] = out["algorithmic_tflops"] / sustained_tflops
    del model, store
    torch.cuda.empty_cache()
    return out
