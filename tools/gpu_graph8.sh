#!/bin/bash
# data-parallel step at N=8: CUDA-graph replay (opt-in) vs eager launches; each must print its line and exit cleanly
cd /root/repo; mkdir -p gpurun_out
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary "$@"; }
RF_BENCH_DP_GRAPH=1 run > gpurun_out/dp8_graph.json 2> gpurun_out/dp8_graph.err; echo "graph rc=$?"
run --no-graph > gpurun_out/dp8_eager.json 2> gpurun_out/dp8_eager.err; echo "eager rc=$?"
for f in gpurun_out/dp8_graph.json gpurun_out/dp8_eager.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['last_loss'], 'launches', d['gpu_launches'], d['config'].get('launch'), d['clocks'])"; done
