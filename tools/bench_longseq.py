"""BASELINE config 5 — long-sequence stress: B=2 x 4096 tokens, attention_window 64 -> 512 sweep, encoder fwd+bwd on one
B200 (longformer-base shape, 12 layers, dropout 0).  Prints one JSON line per window."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_extras import longseq_bench

for r in longseq_bench(torch.device("cuda", 0), steps=5):
    print(json.dumps(dict(r, config="long-sequence stress (BASELINE configs[4])")), flush=True)
