#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "band" -p no:cacheprovider > gpurun_out/t_band.log 2>&1; echo "== band tests exit $?: $(tail -1 gpurun_out/t_band.log)"; grep -E "^E  |Error|FAILED" gpurun_out/t_band.log | head
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/m_all.log 2>&1; echo "== model tests exit $?: $(tail -1 gpurun_out/m_all.log)"; grep -E "^E  |Error|FAILED" gpurun_out/m_all.log | head
timeout 300 python tools/prof_kernels.py attn_fwd attn_bwd attn_bwd_nodrop 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'], d['clocks'])"
