"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, collections, sys
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0     # launches to skip (warm-up steps)
rows = list(csv.reader(open(path, errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); mu = h.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0]); n = 0
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    if r[mu] == "us": v *= 1e3
    elif r[mu] == "ms": v *= 1e6
    n += 1
    if n <= skip: continue
    name = r[kn].split("(")[0][:60]
    d[name][0] += 1; d[name][1] += v
tot = sum(v[1] for v in d.values())
print(f"{'kernel':62s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for k, v in sorted(d.items(), key=lambda x: -x[1][1]):
    print(f"{k:62s} {v[0]:5d} {v[1]/1e3:10.1f} {v[1]/1e3/v[0]:8.1f} {100*v[1]/tot:5.1f}%")
print(f"total {tot/1e6:.3f} ms over {n - skip} launches")
