"""BASELINE config 5 — long-sequence stress: B=2 x 4096 tokens, attention_window 64 -> 512 sweep, encoder
fwd+bwd on one B200 (longformer-base shape, 12 layers, dropout 0).  Prints one JSON line per window."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import recformer_b200 as rb
from oracle import recformer_oracle as O

dev = torch.device("cuda", 0)
B, L = 2, 4096
for window in (64, 128, 256, 512):
    cfg = rb.RecformerConfig(attention_window=[window] * 12, max_token_num=L, hidden_dropout_prob=0.0,
                             attention_probs_dropout_prob=0.0)
    model = rb.RecformerModel(cfg).to(dev).train()
    model.strict_checks = False
    batch = {k: v.to(dev) for k, v in O.make_batch(O.OracleConfig(attention_window=[window] * 12), B, L, seed=1,
                                                   ragged=True).items()}
    def step():
        out = model(**batch).pooler_output
        out.float().square().sum().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # algorithmic FLOPs: dense 14 155 776 + band 4*(window+2)*768 per token-layer, x3 for fwd+bwd
    flops = 3 * (14155776 + 4 * (window + 2) * 768) * 12 * B * L
    print(json.dumps({"config": "long-sequence stress (BASELINE configs[4])", "B": B, "L": L, "attention_window": window,
                      "ms_per_step": ms, "seqs_per_s": B / (ms / 1e3), "tokens_per_s": B * L / (ms / 1e3),
                      "algorithmic_tflops": flops / (ms / 1e3) / 1e12}), flush=True)
    del model
    torch.cuda.empty_cache()
