#!/bin/bash
# data-parallel CUDA-graph step at N=2 (opt-in): must print its line AND exit cleanly
cd /root/repo; mkdir -p gpurun_out
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary "$@"; }
RF_BENCH_DP_GRAPH=1 run > gpurun_out/dp2_graph.json 2> gpurun_out/dp2_graph.err; echo "graph rc=$?"; tail -3 gpurun_out/dp2_graph.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/dp2_graph.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['last_loss'], 'launches', d['gpu_launches'], d['config'].get('launch'))"
