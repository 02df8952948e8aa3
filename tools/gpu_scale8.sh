#!/bin/bash
# 8-GPU validation: bench at N=1 (no extras) and N=8 (full: eval leg, C3, result-equality checks) on the same box
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_n1_8box.log 2> gpurun_out/r02_bench_n1_8box.err; echo "== n1 exit $?"
timeout 900 $TR --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.log 2> gpurun_out/r02_bench_n8.err; echo "== n8 exit $?"; tail -3 gpurun_out/r02_bench_n8.err | cut -c1-300
python - <<'PY'
import json
for f in ("r02_bench_n1_8box", "r02_bench_n8"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        s = d.get("secondary") or {}
        print(f, "train", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], "| eval", round(s.get("value", 0)), "ms", s.get("ms_per_pass"), "e2e", round((s.get("e2e") or {}).get("value", 0)), "kernel frac", (s.get("roofline") or {}).get("frac"))
        print("   checks", json.dumps(d.get("checks"))[:1500]); print("   c3", json.dumps((d.get("extras") or {}).get("pretrain_c3"))[:500])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
