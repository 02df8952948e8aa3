"""GPU timeline of the training step from CUPTI (torch.profiler): per-stream busy time, idle gaps on the
main stream, and which kernels precede the largest gaps.  Diagnostic only -- timings under a profiler are
never bench values.  Usage: python tools/step_timeline.py [steps]"""
import os
import sys
import collections

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from recformer_b200.optim import FusedAdamW

graph = "--graph" in sys.argv
e2e = "--e2e" in sys.argv          # with --graph: pinned host batches in, loss.item() out every step (bench.py's e2e loop)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = int(args[0]) if args else 2
dev = torch.device("cuda", 0)
model, cfg = bench.build_model(dev)
model.train()
opt = FusedAdamW(model, lr=5e-5, weight_decay=0.01)
host, batches = bench.make_batches(2, dev, 0)
for w in range(4):
    bench.train_step(model, opt, batches[w % 2], 1)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
if graph:      # the captured step the bench times (python tools/step_timeline.py 2 --graph)
    from recformer_b200.graph import GraphedTrainStep
    gstep = GraphedTrainStep(model, opt, batches[0])
    for w in range(3):
        gstep(batches[w % 2])
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(steps):
        if graph and e2e:
            float(gstep(host[i % 2]).item())
        elif graph:
            gstep(batches[i % 2])
        else:
            bench.train_step(model, opt, batches[i % 2], 1)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = []
for e in ev:
    tr = e.time_range
    ks.append((tr.start, tr.end, e.name, getattr(e, "device_index", 0), getattr(e, "stream", None)))
ks.sort()
if "--dump" in sys.argv:     # every device event (start us, duration us, name) for offline reading
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/timeline_events.tsv", "w") as f:
        for s_, e_, n_, *_ in ks:
            f.write(f"{s_ - ks[0][0]:.2f}\t{e_ - s_:.2f}\t{n_[:90]}\n")
t0, t1 = ks[0][0], max(k[1] for k in ks)
print(f"{len(ks)} device events over {(t1 - t0) / 1e3:.3f} ms ({steps} steps -> {(t1 - t0) / 1e3 / steps:.3f} ms/step)")
# union busy time
busy, cur_s, cur_e = 0.0, None, None
for s, e, *_ in ks:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"union busy {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms ({100 * (1 - busy / (t1 - t0)):.1f}%)")
by_stream = collections.defaultdict(list)
for k in ks:
    by_stream[k[4]].append(k)
for st, lst in by_stream.items():
    tot = sum(e - s for s, e, *_ in lst)
    print(f"stream {st}: {len(lst)} events, busy {tot / 1e3:.3f} ms")
main = max(by_stream.values(), key=len)
gaps = []
for a, b in zip(main, main[1:]):
    gaps.append((b[0] - a[1], a[2][:50], b[2][:50]))
g = [x[0] for x in gaps]
print(f"main stream: {len(g)} gaps, total {sum(g) / 1e3:.3f} ms, median {sorted(g)[len(g) // 2]:.2f} us")
hist = collections.Counter(min(int(x), 20) for x in g)
print("gap histogram (us -> count):", sorted(hist.items()))
agg = collections.defaultdict(lambda: [0, 0.0])
for d, a, b in gaps:
    agg[(a, b)][0] += 1
    agg[(a, b)][1] += d
print("largest gap totals (prev -> next):")
for (a, b), (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {tot / steps:8.1f} us/step  n={n // steps:3d}  avg {tot / n:6.2f}  {a}  ->  {b}")
dur = collections.defaultdict(lambda: [0, 0.0])
for s, e, n, *_ in ks:
    dur[n[:60]][0] += 1
    dur[n[:60]][1] += e - s
print("kernel totals per step (warm, overlapped):")
for n, (c, tot) in sorted(dur.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {tot / steps:8.1f} us  n={c // steps:3d}  avg {tot / c:7.2f}  {n}")
