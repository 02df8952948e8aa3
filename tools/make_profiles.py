"""Summarise gpurun_out/<round>_* profiling artefacts into tracked text files under profiles/.

    python tools/make_profiles.py r01
Writes, per round:
  profiles/<r>_launches_bench.csv        the ncu launch list of `bench.py --steps 1 --warmup 3 --no-graph` (as captured)
  profiles/<r>_launches_summary.txt      per-kernel totals / shares of that list (cold-cache, serialised)
  profiles/<r>_kernel_times.txt          CUDA-event timings of each hot kernel at the C2 shapes (no profiler)
  profiles/<r>_ncu_<set>.txt             roofline counters + stall reasons + SASS hot spots of the --set full captures
  profiles/<r>_bench_plain.json          the bench line of the un-profiled run made in the same call
"""
import csv, os, shutil, subprocess, sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max ", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg ", "sm__cycles_elapsed.avg ", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct", "sm__inst_executed_pipe_tensor", "smsp__inst_executed.sum ",
        "smsp__issue_active.avg.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum ",
        "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak", "smsp__average_warps_issue_stalled"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summarise_rep(rep, dst, topn=12):
    raw = ncu_csv(rep, "raw")
    src = ncu_csv(rep, "source")
    lines = [f"# {os.path.basename(rep)}  (ncu --set full --clock-control none; per-launch values)", ""]
    h, units = raw[0], raw[1]
    hot = {}
    i = 0
    while i < len(src):                      # per-kernel SASS hot spots
        if src[i] and src[i][0] == "Kernel Name":
            name, hh = src[i][1], src[i + 1]
            isrc, ismp = hh.index("Source"), hh.index("# Samples")
            stall = [(k, c) for k, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
            data, j = [], i + 2
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                r = src[j]
                if len(r) >= len(hh):
                    try: data.append((int(r[ismp]), r))
                    except ValueError: pass
                j += 1
            tot = sum(n for n, _ in data) or 1
            agg = {}
            for n, r in data:
                for k, c in stall:
                    if r[k] not in ("", "0"): agg[c] = agg.get(c, 0) + int(r[k])
            rows = [f"  stall samples {tot}: " + ", ".join(f"{c[6:]} {100*v/tot:.0f}%" for c, v in sorted(agg.items(), key=lambda x: -x[1])[:7])]
            for n, r in sorted(data, key=lambda x: -x[0])[:topn]:
                st = sorted(((c[6:], int(r[k])) for k, c in stall if r[k] not in ("", "0")), key=lambda x: -x[1])[:2]
                rows.append(f"  {100*n/tot:5.1f}%  {r[isrc][:70]:70s} {st}")
            hot.setdefault(name, []).append(rows)
            i = j
        else:
            i += 1
    seen = {}
    for r in raw[2:]:
        name = r[h.index("Kernel Name")]
        k = seen.get(name, 0); seen[name] = k + 1
        lines.append("=" * 100); lines.append(f"{name}   [capture #{k}]")
        for idx, c in enumerate(h):
            cc = c + " "
            if any(key in cc for key in KEYS) and r[idx] not in ("", "n/a") and "not_issued" not in c and \
               "pct_of_peak_sustained_active" not in c.replace("warps_active", "").replace("issue_active", ""):
                lines.append(f"  {c:84s} {units[idx]:12s} {r[idx][:40]}")
        # tcgen05 tensor-pipe utilisation.  `sm__pipe_tensor_cycles_active_realtime...pct` under-reports UTCHMMA work
        # by ~5x on sm_100a (it read 3.5-18 % on GEMMs that CUDA events put at 0.86-1.34 PFLOP/s).  The sub-pipe cycle
        # counter is exact: it equals (UTCHMMA instructions per SM) x (cycles per instruction at the dense bf16 rate,
        # 8192 FLOP/clk/SM) x 4 (it aggregates the SM's four sub-partitions), so active share = counter / 4 / elapsed.
        try:
            hm = float(r[h.index("TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg")])
            el = float(r[h.index("sm__cycles_elapsed.max")])
            ghz = float(r[h.index("sm__cycles_elapsed.max.per_second")])
            lines.append(f"  {'DERIVED tensor pipe (UTCHMMA) active = hmma_cycles_active_realtime.avg / 4 / sm__cycles_elapsed.max':84s} "
                         f"{'%':12s} {100 * hm / 4 / el:.1f}   (= {hm / 4 * 8192 * 148 * ghz / el / 1e3:.0f} TFLOP/s at {ghz:.2f} GHz)")
        except (ValueError, IndexError):
            pass
        norm = lambda n: n.replace("rf::", "").replace("void ", "").replace("(bool)", "").replace("(int)", "").split("(")[0].replace(" ", "")
        hs = []
        for key, val in hot.items():
            if norm(key) == norm(name):
                hs = val
        if k < len(hs):
            lines.append("  -- SASS hot spots (share of stall samples, instruction, top stall reasons)")
            lines += hs[k]
    open(dst, "w").write("\n".join(lines) + "\n")


def summarise_launches(path, dst):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg, n = {}, 0
    for r in rows[hdr + 1:]:
        if len(r) <= mv: continue
        try: v = float(r[mv].replace(",", ""))
        except ValueError: continue
        v *= {"us": 1e3, "ms": 1e6}.get(r[mu], 1.0)
        n += 1
        a = agg.setdefault(r[kn].split("(")[0][:64], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# per-kernel totals of {os.path.basename(path)}: ncu --metrics gpu__time_duration.sum --clock-control none on",
           "# `bench.py --steps 1 --warmup 3 --no-graph` (4 training steps + eval leg); cold-cache, serialised launch times",
           f"{'kernel':66s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{k:66s} {v[0]:5d} {v[1]/1e3:10.1f} {v[1]/1e3/v[0]:8.1f} {100*v[1]/tot:5.1f}%")
    out.append(f"total {tot/1e6:.3f} ms over {n} launches")
    open(dst, "w").write("\n".join(out) + "\n")


for f in os.listdir(G):
    if not f.startswith(R + "_"): continue
    src = os.path.join(G, f)
    if f.endswith(".ncu-rep"):
        summarise_rep(src, os.path.join(P, f.replace(".ncu-rep", ".txt")))
    elif f.endswith("_launches_bench.csv"):
        shutil.copy(src, os.path.join(P, f))
        summarise_launches(src, os.path.join(P, f.replace("_bench.csv", "_summary.txt")))
    elif f.endswith("_kernel_times.log"):
        shutil.copy(src, os.path.join(P, f.replace(".log", ".txt")))
    elif f.endswith("_bench_plain.log"):
        line = [l for l in open(src).read().splitlines() if l.startswith("{")]
        if line: open(os.path.join(P, f.replace(".log", ".json")), "w").write(line[-1] + "\n")
# per-kernel DRAM traffic (bench.py reports it as roofline.traffic)
import json, re
traffic = {}
for f in sorted(os.listdir(G)):
    if f.startswith(R + "_") and f.endswith(".ncu-rep"):
        raw = ncu_csv(os.path.join(G, f), "raw")
        h = raw[0]
        ik, ir, iw, it = (h.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        ur, uw, ut = raw[1][ir], raw[1][iw], raw[1][it]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "second": 1e6}
        for r in raw[2:]:
            name = re.sub(r"\(CUtensorMap.*", "", r[ik])
            if "shard" in f:          # a capture at another problem size must not be averaged into the full-size entry
                name += " [125k-item shard]"
            t = traffic.setdefault(name, {"dram_bytes": 0.0, "time_us": 0.0, "n": 0, "source": f})
            t["dram_bytes"] += float(r[ir]) * mult[ur] + float(r[iw]) * mult[uw]
            t["time_us"] += float(r[it]) * mult[ut]
            t["n"] += 1
for t in traffic.values():
    t["dram_bytes"] /= t["n"]; t["time_us"] /= t["n"]
json.dump(traffic, open(os.path.join(P, R + "_traffic.json"), "w"), indent=1, sort_keys=True)
# SASS evidence: tcgen05 / TMEM / TMA mnemonics per kernel of the shipped library (cuobjdump, no GPU needed)
lib = os.path.join(ROOT, "recformer_b200", "librecformer_b200.so")
if os.path.exists(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, cur = {}, None
    pats = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA.", "RED.E", "ATOM")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0][:90]
            counts[cur] = dict.fromkeys(pats, 0)
        elif cur:
            for pat in pats:
                if pat in line: counts[cur][pat] += 1
    out = [f"# cuobjdump -sass recformer_b200/librecformer_b200.so: instruction counts per kernel ({R}); UTCHMMA = tcgen05.mma,",
           "# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, UTCBAR = tcgen05.commit, HMMA. = legacy mma.sync (none expected)",
           f"{'kernel':92s} " + " ".join(f"{x:>8s}" for x in pats)]
    for k, v in sorted(counts.items()):
        out.append(f"{k:92s} " + " ".join(f"{v[x]:8d}" for x in pats))
    tot = {x: sum(v[x] for v in counts.values()) for x in pats}
    out.append(f"{'TOTAL':92s} " + " ".join(f"{tot[x]:8d}" for x in pats))
    open(os.path.join(P, R + "_sass_counts.txt"), "w").write("\n".join(out) + "\n")
print(sorted(os.listdir(P)))
