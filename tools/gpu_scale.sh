#!/bin/bash
# scaling check on an 8-GPU box: N = 8 and 4 (N = 1, 2 are measured on smaller boxes)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err
  echo "== N=$n exit $?"; tail -1 gpurun_out/bench_n$n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['secondary'].get('value'), d['secondary'].get('ms_per_pass'), d['secondary'].get('error'))"
done
