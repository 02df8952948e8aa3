#!/bin/bash
# Run the per-kernel GPU parity tests group by group, each in its own process with a timeout so a
# faulting kernel cannot poison (or hang) the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rc=0
for grp in gemm_tn gemm_q gemm_gelu gemm_dgrad gemm_wgrad gemm_dropout prepare layernorm band_attention global_attention normalize topk_sharded cosine_ce cast; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$grp" -p no:cacheprovider > gpurun_out/t_$grp.log 2>&1
  c=$?
  echo "== $grp exit $c: $(tail -1 gpurun_out/t_$grp.log)"
  if [ $c -ne 0 ]; then rc=1; grep -E "^E  |Error|error" gpurun_out/t_$grp.log | head -12; fi
done
exit $rc
