#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or topk or logits or cosine" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1; echo "== gemm/score tests exit $?: $(tail -1 gpurun_out/t_gemm.log)"; grep -E "^E  |Error|FAILED" gpurun_out/t_gemm.log | head
timeout 300 python tools/prof_kernels.py gemm_qkv gemm_up_gelu gemm_down_res gemm_dgrad_dgelu gemm_wgrad_up gemm_wgrad_qkv score_topk 2>&1 | tail -7
echo NOEPI; RF_DEBUG_GEMM_NOEPI=1 timeout 300 python tools/prof_kernels.py gemm_qkv gemm_up_gelu gemm_wgrad_up 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'], d['clocks'])"
