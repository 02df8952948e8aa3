#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29561 tools/bench_allreduce.py 2>&1 | grep -E "allreduce_bench" | cut -c1-1500
B="bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-secondary"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], d["config"]["launch"][:12], "gemm_ms", round(d["roofline"]["gemm_ms_per_step"],2))'
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-secondary 2>/dev/null | python -c "$P" n1
timeout 900 $TR --nproc-per-node 2 --master-port 29541 $B 2>/dev/null | python -c "$P" n2-buckets4
RF_DP_SYNC_AT_END=1 timeout 900 $TR --nproc-per-node 2 --master-port 29542 $B 2>/dev/null | python -c "$P" n2-sync-at-end
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-secondary 2>/dev/null | python -c "$P" n1
timeout 900 $TR --nproc-per-node 2 --master-port 29543 $B 2>/dev/null | python -c "$P" n2-buckets4
