#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest gpu exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^E  |Error|FAILED" gpurun_out/pytest_gpu.log | head
timeout 300 python tools/prof_kernels.py gemm_qkv gemm_up_gelu gemm_down_res gemm_dgrad_dgelu gemm_wgrad_up gemm_wgrad_qkv attn_fwd attn_bwd global_fwd global_bwd ln_fwd ln_bwd score_topk 2>&1 | tail -13
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'], d['clocks']); print(d['secondary']['value'], d['secondary']['roofline']['frac'])"
