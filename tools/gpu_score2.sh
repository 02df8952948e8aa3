#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "topk or logits or cosine or band" -p no:cacheprovider > gpurun_out/t_score.log 2>&1; echo "== score/band tests exit $?: $(tail -1 gpurun_out/t_score.log)"; grep -E "^E  |Error|FAILED" gpurun_out/t_score.log | head
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/m_all.log 2>&1; echo "== model tests exit $?: $(tail -1 gpurun_out/m_all.log)"
timeout 300 python tools/prof_kernels.py score_topk attn_bwd 2>&1 | tail -2
RF_PROF_ITEMS=125000 timeout 300 python tools/prof_kernels.py score_topk 2>&1 | tail -1
RF_PROF_ITEMS=250000 timeout 300 python tools/prof_kernels.py score_topk 2>&1 | tail -1
