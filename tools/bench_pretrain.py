"""BASELINE config 3 — pretraining step: MLM + item-item contrastive with in-batch negatives, batch 64 per GPU,
data-parallel (launch with torchrun for N > 1).  One step = 4 encoder passes (history a <= 1024 tokens, target item
b <= 128 tokens, and their masked copies) + LM head on the masked rows + contrastive CE + backward + AdamW.
Prints one JSON line (rank 0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import recformer_b200 as rb
from recformer_b200 import dist as rdist
from recformer_b200.optim import FusedAdamW
from oracle import recformer_oracle as O      # synthetic weights / batches only

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, LA, LB = int(os.environ.get("RF_PRETRAIN_B", "64")), 1024, 128
steps, warmup = int(os.environ.get("RF_STEPS", "5")), 3
ocfg = O.OracleConfig()
cfg = rb.RecformerConfig(attention_window=[64] * 12, max_token_num=LA, max_item_embeddings=51, max_attr_num=3, max_attr_length=32)
model = rb.RecformerForPretraining(cfg)
model.load_state_dict(O.make_pretrain_state_dict(ocfg, seed=0), strict=True)
model = model.to(dev).train()
model.longformer.strict_checks = False
opt = FusedAdamW(model, lr=5e-5, weight_decay=0.01)
head_opt = torch.optim.AdamW(model.lm_head.parameters(), lr=5e-5, weight_decay=0.01, fused=True)
sync = rdist.GradSync(model, passes_per_step=4) if world > 1 else None
batches = [{k: v.to(dev) for k, v in O.make_pretrain_batch(ocfg, B, LA, LB, seed=100 * rank + i).items()} for i in range(2)]


def step(i):
    out = model(**batches[i % 2])
    opt.zero_grad(); head_opt.zero_grad(set_to_none=True)
    out.loss.backward()
    if sync is not None:
        sync.finish()
        for p in model.lm_head.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad)
                p.grad.mul_(1.0 / world)
    opt.step(grad_scale=1.0 / world)
    head_opt.step()
    return out


for i in range(warmup):
    step(i)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    out = step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
if world > 1:
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    tokens = B * (2 * LA + 2 * LB)     # padded tokens through the encoder per step and rank
    flops = 3 * (14155776 + 4 * 66 * 768) * 12 * tokens
    print(json.dumps({"config": "pretraining step (BASELINE configs[2]): MLM + in-batch contrastive, 4 encoder passes, "
                                f"batch {B}/GPU x (1024 + 128) tokens, dropout 0.1, fwd+bwd+AdamW",
                      "n_gpus": world, "ms_per_step": ms, "pairs_per_s": world * B / (ms / 1e3),
                      "encoder_tokens_per_s": world * tokens / (ms / 1e3),
                      "encoder_algorithmic_tflops_per_gpu": flops / (ms / 1e3) / 1e12, "last_loss": float(out.loss.item()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
