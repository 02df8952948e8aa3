#!/bin/bash
# 2-GPU validation: equality checks (incl. the peer-memory top-k exchange), eval leg with both exchanges
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 tools/check_multigpu.py --full > gpurun_out/check_n2.log 2>&1; echo "== check_multigpu exit $?"; grep check_multigpu gpurun_out/check_n2.log | cut -c1-2200; grep -E "Error|error" gpurun_out/check_n2.log | head -5
run() { name=$1; shift
  env "$@" timeout 900 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/dp2_$name.log 2> gpurun_out/dp2_$name.err; echo "== bench $name exit $?"; tail -2 gpurun_out/dp2_$name.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/dp2_$name.log") if l.startswith("{")][-1])
    s = d["secondary"]
    print("$name", "train", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "| eval", round(s["value"]), "ms/pass", round(s["ms_per_pass"], 4), "e2e", round(s["e2e"]["value"]), "launches", s["gpu_launches"], s["metrics"])
    print("   ", s["config"]["workload"][-120:])
except Exception as ex:
    print("$name unreadable", ex)
PY
}
run peer A=1
run nccl RF_BENCH_NCCL_TOPK=1
