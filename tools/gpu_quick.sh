#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -8 | cut -c1-300
timeout 300 python tools/prof_kernels.py global_fwd global_bwd 2>&1 | tail -2
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], "gemm_ms", round(d["roofline"]["gemm_ms_per_step"],2), "eager", round(d["roofline"]["eager_step_ms"],2))'
B="bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-secondary"
timeout 600 python $B 2>/dev/null | python -c "$P" normal
RF_DEBUG_SKIP_GLOBAL=1 timeout 600 python $B 2>/dev/null | python -c "$P" skip-global
RF_DEBUG_NO_OVERLAP=1 timeout 600 python $B 2>/dev/null | python -c "$P" no-overlap
timeout 600 python $B 2>/dev/null | python -c "$P" normal
