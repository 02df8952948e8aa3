#!/bin/bash
# attention iteration loop: parity tests of the band kernels + dropout-mask tests, then CUDA-event timings (+ optional ncu)
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -p no:cacheprovider --tb=short -k "band or dropout or mask" > gpurun_out/t_attn.log 2>&1; echo "== attn tests exit $?: $(tail -1 gpurun_out/t_attn.log)"; grep -E "^E  |FAILED|^ERROR" gpurun_out/t_attn.log | head -30
timeout 300 python tools/prof_kernels.py attn_fwd attn_fwd_nodrop attn_bwd attn_bwd_nodrop 2>&1 | tee gpurun_out/attn_times.log
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"band_attn_bwd" -s 1 -c 1 -o gpurun_out/dbg_ncu_attn_bwd python tools/prof_kernels.py attn_fwd attn_bwd > /dev/null 2>&1; echo "ncu attn bwd exit $?"
fi
