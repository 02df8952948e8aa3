#!/bin/bash
# CUDA-graph step: tests + A/B bench
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "graphed or dropout" > gpurun_out/graph_tests.log 2>&1
echo "graph tests rc=$?" >> gpurun_out/graph_tests.log
tail -30 gpurun_out/graph_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/graph_bench.json 2> gpurun_out/graph_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/graph_bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-graph > gpurun_out/graph_bench_eager.json 2>> gpurun_out/graph_bench.err
for f in gpurun_out/graph_bench.json gpurun_out/graph_bench_eager.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['last_loss'], 'launches', d['gpu_launches'], d['config'].get('launch'))"; done
