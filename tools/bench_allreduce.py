"""Micro-benchmark (torchrun, N GPUs): NCCL all_reduce vs torch symmetric-memory multimem / two-shot all-reduce of the
data-parallel gradient payload (149 M bf16 elements = 298 MB), device-timed, max over ranks."""
import json, os, sys
import torch, torch.distributed as dist
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 149 * 1024 * 1024
res = {}
def timed(fn, iters=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
x = torch.ones(n, dtype=torch.bfloat16, device=dev)
res["nccl_bf16_298MB_ms"] = timed(lambda: dist.all_reduce(x))
x32 = torch.ones(n, dtype=torch.float32, device=dev)
res["nccl_fp32_596MB_ms"] = timed(lambda: dist.all_reduce(x32))
chunks = [x[i * (n // 26):(i + 1) * (n // 26)] for i in range(26)]
res["nccl_bf16_26_chunks_ms"] = timed(lambda: [dist.all_reduce(c) for c in chunks])
try:
    import torch.distributed._symmetric_memory as symm_mem
    gname = dist.group.WORLD.group_name
    s = symm_mem.empty(n, dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(s, dist.group.WORLD)
    s.fill_(1.0)
    res["symm_multicast_ptr"] = int(getattr(hdl, "multicast_ptr", 0) or 0) != 0
    try:
        res["symm_multimem_298MB_ms"] = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(s, "sum", gname))
        s.fill_(1.0); torch.ops.symm_mem.multimem_all_reduce_(s, "sum", gname); torch.cuda.synchronize()
        res["symm_multimem_correct"] = bool((s[:1000].float() == world).all().item() and (s[-1000:].float() == world).all().item())
        sl = [s[i * (n // 26):(i + 1) * (n // 26)] for i in range(26)]
        res["symm_multimem_26_slices_ms"] = timed(lambda: [torch.ops.symm_mem.multimem_all_reduce_(c, "sum", gname) for c in sl])
    except Exception as ex:
        res["symm_multimem_error"] = repr(ex)[:300]
    try:
        res["symm_two_shot_298MB_ms"] = timed(lambda: torch.ops.symm_mem.two_shot_all_reduce_(s, "sum", gname))
    except Exception as ex:
        res["symm_two_shot_error"] = repr(ex)[:300]
except Exception as ex:
    res["symm_error"] = repr(ex)[:300]
if rank == 0: print(json.dumps({"allreduce_bench": res, "world": world}), flush=True)
dist.barrier(); dist.destroy_process_group()
