"""A/B timing of the GEMM shapes of a training step: 30 back-to-back launches per shape between two CUDA events
(RF_LIB_PATH selects the build).  M = rows processed for a typical ragged batch (57 of 64 row tiles) unless AB_M is set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recformer_b200 import ops
dev = "cuda"
M = int(os.environ.get("AB_M", 57 * 256)); E, F = 768, 3072
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(torch.bfloat16)
rf = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
x, Wqkv, bqkv = rb(M, E), rb(3 * E, E, sc=0.02), rf(3 * E, sc=0.02)
Wo, bo = rb(E, E, sc=0.02), rf(E, sc=0.02)
W1, b1, W2, b2 = rb(F, E, sc=0.02), rf(F, sc=0.02), rb(E, F, sc=0.02), rf(E, sc=0.02)
qkv = torch.empty(M, 3 * E, dtype=torch.bfloat16, device=dev)
u = torch.empty(M, F, dtype=torch.bfloat16, device=dev); gl = torch.empty_like(u)
res32 = rf(M, E); pre = torch.empty(M, E, dtype=torch.float32, device=dev)
dY = rb(M, E, sc=0.01); dU = torch.empty(M, F, dtype=torch.bfloat16, device=dev)
dx = torch.empty(M, E, dtype=torch.bfloat16, device=dev); resb = rb(M, E)
dW1 = torch.zeros(F, E, dtype=torch.float32, device=dev)
CASES = {
    "qkv": (lambda: ops.gemm(x, Wqkv, out=qkv, bias=bqkv, scale=0.125, scale_ncols=E), 3 * E * E),
    "wo_res_drop": (lambda: ops.gemm(x, Wo, out=pre, bias=bo, residual=res32, drop_p=0.1, drop_seed=1), E * E),
    "wo_res_nodrop": (lambda: ops.gemm(x, Wo, out=pre, bias=bo, residual=res32), E * E),
    "wo_bf16_plain": (lambda: ops.gemm(x, Wo, out=dx, bias=bo), E * E),
    "down_res_nodrop": (lambda: ops.gemm(gl, W2, out=pre, bias=b2, residual=res32), F * E),
    "up_gelu": (lambda: ops.gemm(x, W1, out=u, bias=b1, epi=ops.EPI_GELU, out2=gl), F * E),
    "down_res_drop": (lambda: ops.gemm(gl, W2, out=pre, bias=b2, residual=res32, drop_p=0.1, drop_seed=1), F * E),
    "dgrad_down_dgelu": (lambda: ops.gemm(dY, W2, out=dU, b_mn_major=True, epi=ops.EPI_DGELU, aux=u), F * E),
    "dgrad_up": (lambda: ops.gemm(dU, W1, out=dx, b_mn_major=True, residual=resb), F * E),
    "dgrad_wo": (lambda: ops.gemm(dY, Wo, out=dx, b_mn_major=True), E * E),
    "dgrad_qkv": (lambda: ops.gemm(qkv, Wqkv, out=dx, b_mn_major=True, residual=resb), 3 * E * E),
    "wgrad_up": (lambda: ops.gemm(dU, x, out=dW1, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=2), F * E),
}
tot = 0.0
for name, (fn, kn) in CASES.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    tot += us
    print(f"{name:18s} {us:7.1f} us  {2 * M * kn / us / 1e6:7.1f} TFLOP/s")
print(f"sum {tot:.1f} us  (M = {M})")
