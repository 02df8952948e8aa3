"""In-kernel phase clocks of the band attention kernels (build with RF_NVCC_DEFINES="-DRF_KTIMING").
Prints, per phase, the cycles between consecutive clock64() stamps of thread 0 (see the KT(k) marks in
csrc/attention_fwd.cu / attention_bwd.cu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recformer_b200 import ops, _lib

dev = "cuda"
B, L, H, E = 16, 1024, 12, 768
T = B * L
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).to(torch.bfloat16)
qkv = rb(T, 3 * E)
mask = torch.ones(B, L, dtype=torch.uint8, device=dev); mask[:, 0] = 2
for b in range(1, B):
    mask[b, L - 31 * b:] = 0
ctx = torch.empty(T, E, dtype=torch.bfloat16, device=dev); lse = torch.empty(B, H, L, device=dev)
dctx = rb(T, E, sc=0.01); dqkv = torch.empty(T, 3 * E, dtype=torch.bfloat16, device=dev)
scratch = torch.empty(T, 2 * E, dtype=torch.float32, device=dev)
drop = float(os.environ.get("KT_DROP", "0.1"))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
kb = ops.band_attn_keepbits(B, L, H, dev) if os.environ.get("KT_KEEPBITS", "1") == "1" else None
lib = _lib.lib()
for name, fn, sym in (("fwd", lambda: ops.band_attn_fwd(qkv, mask, B, L, H, 32, ctx=ctx, lse=lse, drop_p=drop, drop_seed=3, keepbits=kb), "rf_debug_ktiming_fwd"),
                      ("bwd", lambda: ops.band_attn_bwd(qkv, mask, B, L, H, 32, ctx, lse, dctx, dqkv, scratch, drop_p=drop, drop_seed=3, keepbits=kb), "rf_debug_ktiming_bwd")):
    for _ in range(2):
        flush.zero_(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"== {name}: {e0.elapsed_time(e1) * 1e3:.1f} us (dropout {drop})")
    if True:
        buf = (C.c_longlong * (4 * 16 * 16))()
        f = getattr(lib, sym)
        f.argtypes = [C.c_void_p]; f.restype = C.c_int
        assert f(C.cast(buf, C.c_void_p)) == 0
        order = [0, 12, 1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 10, 11] if name == "bwd" else list(range(11))
        base = [buf[i * 16 + 0] for i in range(16)]       # warp 0 stamp 0 of each tile
        for w in range(4):
            print(f"  warp {w * 5}: stamps relative to warp 0's tile start, order {order}")
            for i in range(2, 9):
                if base[i] == 0:
                    continue
                print(f"   tile {i:2d}: " + " ".join(f"{buf[(w * 16 + i) * 16 + k] - base[i]:6d}" if buf[(w * 16 + i) * 16 + k] else "     -" for k in order))
        continue
    buf = (C.c_longlong * (16 * 12))()
    f = getattr(lib, sym)
    f.argtypes = [C.c_void_p]; f.restype = C.c_int
    assert f(C.cast(buf, C.c_void_p)) == 0
    rows = [[buf[i * 12 + k] for k in range(12)] for i in range(16)]
    for i, r in enumerate(rows):
        if r[0] == 0:
            continue
        d = [r[k + 1] - r[k] for k in range(11)]
        print(f"  {i:2d} total {r[11] - r[0]:6d} | " + " ".join(f"{x:6d}" for x in d))
    if name == "bwd":
        it = [rows[i + 1][0] - rows[i][0] for i in range(15) if rows[i + 1][0] and rows[i][0]]
        print("  tile period:", it)
