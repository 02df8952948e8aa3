"""BASELINE config 3 — pretraining step (launch with torchrun for N > 1); prints one JSON line (rank 0).
RF_PRETRAIN_B / RF_STEPS override the per-GPU batch and the timed steps."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tools.bench_extras import pretrain_bench

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
res = pretrain_bench(dev, rank, world, steps=int(os.environ.get("RF_STEPS", "5")), B=int(os.environ.get("RF_PRETRAIN_B", "64")))
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
