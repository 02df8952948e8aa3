#!/bin/bash
# 8-GPU box: full bench at N=8 (train + sharded eval), NCCL channel sweep on the training leg, pretraining at N=8
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 600 env python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
echo "== N=8 exit $?"; tail -1 gpurun_out/bench_n8.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm', d['roofline']['achieved'], 'eval', d['secondary'].get('value'), d['secondary'].get('ms_per_pass'), d['secondary'].get('error'))"
for ch in 4 8 16; do
  NCCL_MAX_NCHANNELS=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$ch bench.py --gpus 8 --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench_n8_ch$ch.log 2>/dev/null
  echo "NCCL_MAX_NCHANNELS=$ch"; tail -1 gpurun_out/bench_n8_ch$ch.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  train', d['value'], d['ms_per_step'], 'gemm', d['roofline']['achieved'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29640 tools/bench_pretrain.py 2>/dev/null | tail -1 | tee gpurun_out/r01_pretrain_n8.jsonl
