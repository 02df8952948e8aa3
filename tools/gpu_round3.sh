#!/bin/bash
mkdir -p gpurun_out
rc=0
for grp in layernorm prepare gemm_q; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$grp" -p no:cacheprovider > gpurun_out/t_$grp.log 2>&1
  c=$?; echo "== $grp exit $c: $(tail -1 gpurun_out/t_$grp.log)"
  if [ $c -ne 0 ]; then rc=1; grep -E "^E  |Error|error" gpurun_out/t_$grp.log | head -12; fi
done
for grp in forward_matches recall_ndcg train_step train_gradients dropout_training; do
  timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "$grp" -p no:cacheprovider > gpurun_out/m_$grp.log 2>&1
  c=$?; echo "== $grp exit $c: $(tail -1 gpurun_out/m_$grp.log)"
  grep -E "hidden .* pooled|worst relative" gpurun_out/m_$grp.log
  if [ $c -ne 0 ]; then rc=1; grep -E "^E  |Error|error" gpurun_out/m_$grp.log | head -20; fi
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
exit $rc
