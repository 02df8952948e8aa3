#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "topk or logits or cosine" -p no:cacheprovider > gpurun_out/t_score.log 2>&1; echo "== score tests exit $?: $(tail -1 gpurun_out/t_score.log)"; grep -E "^E  |Error|FAILED" gpurun_out/t_score.log | head
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "recall" -p no:cacheprovider > gpurun_out/m_recall.log 2>&1; echo "== recall test exit $?: $(tail -1 gpurun_out/m_recall.log)"
timeout 300 python tools/prof_kernels.py score_topk 2>&1 | tail -3; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"secondary\"])"
