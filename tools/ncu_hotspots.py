"""Top stall locations (SASS) per kernel of an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:90]; h = rows[i + 1]
        isrc = h.index("Source"); ismp = h.index("# Samples"); iex = h.index("Instructions Executed")
        stall = [(k, c) for k, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        data = []; j = i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            r = rows[j]
            if len(r) >= len(h):
                try: data.append((int(r[ismp]), r))
                except ValueError: pass
            j += 1
        tot = sum(n for n, _ in data) or 1
        agg = {}
        for n, r in data:
            for k, c in stall:
                if r[k] not in ("", "0"): agg[c] = agg.get(c, 0) + int(r[k])
        print("=" * 110); print(name); print("samples", tot, "by reason:", sorted(agg.items(), key=lambda x: -x[1])[:8])
        for idx, (n, r) in sorted(enumerate(data), key=lambda x: -x[1][0])[:topn]:
            st = sorted(((c, int(r[k])) for k, c in stall if r[k] not in ("", "0")), key=lambda x: -x[1])[:2]
            print(f"{idx:5d} {100*n/tot:5.1f}% ex={r[iex]:>8s} {r[isrc][:64]:64s} {st}")
        i = j
    else:
        i += 1
