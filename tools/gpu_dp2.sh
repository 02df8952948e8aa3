#!/bin/bash
# 2-GPU validation of the data-parallel step: equality checks, then the bench with and without the bucket-overlapped update
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 tools/check_multigpu.py > gpurun_out/check_n2.log 2>&1; echo "== check_multigpu exit $?"; grep check_multigpu gpurun_out/check_n2.log | cut -c1-1800; grep -E "Error|error" gpurun_out/check_n2.log | head -5
run() { name=$1; shift
  env "$@" timeout 900 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-secondary > gpurun_out/dp2_$name.log 2> gpurun_out/dp2_$name.err; echo "== bench $name exit $?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/dp2_$name.log") if l.startswith("{")][-1])
    print("$name", "train", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"])
except Exception as ex:
    print("$name unreadable", ex)
PY
}
run overlap A=1
run plain RF_DP_PLAIN=1
run overlap2 A=1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --no-secondary > gpurun_out/dp2_n1.log 2> gpurun_out/dp2_n1.err; python -c "
import json
d = json.loads([l for l in open('gpurun_out/dp2_n1.log') if l.startswith('{')][-1]); print('n1', round(d['value'],1), round(d['ms_per_step'],3))"
