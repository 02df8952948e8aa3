#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --tb=short -x -k "padding_tile or 12layer or graph" 2>&1 | tail -15 | cut -c1-400
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-secondary > gpurun_out/q_$name.log 2> gpurun_out/q_$name.err; tail -1 gpurun_out/q_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), d['e2e'].get('last_loss'))" || tail -5 gpurun_out/q_$name.err
}
run skip A=1
run noskip RF_NO_TILE_SKIP=1
