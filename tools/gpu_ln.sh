#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "layernorm or embed" > gpurun_out/ln_tests.log 2>&1; echo "kernel tests rc=$? $(tail -1 gpurun_out/ln_tests.log)"
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "train or graph or pretrain" > gpurun_out/ln_model.log 2>&1; echo "model tests rc=$? $(tail -1 gpurun_out/ln_model.log)"
timeout 300 python tools/prof_kernels.py ln_fwd ln_bwd attn_bwd 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
