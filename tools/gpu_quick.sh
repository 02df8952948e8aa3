#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-secondary > gpurun_out/q_$name.log 2> gpurun_out/q_$name.err; tail -1 gpurun_out/q_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), d['roofline'].get('rows_processed',{}).get('profiled_batch'))" || tail -5 gpurun_out/q_$name.err
}
run skip A=1
run skip_wgaux RF_WGRAD_AUX=1
run skip2 A=1
run skip_wgaux2 RF_WGRAD_AUX=1
