#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "band" -p no:cacheprovider > gpurun_out/t_band.log 2>&1; echo "== band tests exit $?: $(tail -1 gpurun_out/t_band.log)"; grep -E "^E  .*(assert|Error)|FAILED" gpurun_out/t_band.log | head -20
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/m_all.log 2>&1; echo "== model tests exit $?: $(tail -1 gpurun_out/m_all.log)"; grep -E "^E  |FAILED" gpurun_out/m_all.log | head
timeout 300 python tools/prof_kernels.py attn_fwd attn_bwd 2>&1 | tail -2
