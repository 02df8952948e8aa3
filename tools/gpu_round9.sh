#!/bin/bash
mkdir -p gpurun_out
rc=0
for grp in gemm band_attention_bwd; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider > gpurun_out/t_$grp.log 2>&1
  c=$?; echo "== $grp exit $c: $(tail -1 gpurun_out/t_$grp.log)"
  if [ $c -ne 0 ]; then rc=1; grep -E "^E  |Error|error|FAILED" gpurun_out/t_$grp.log | head -20; fi
done
for grp in train_step forward_matches train_gradients dropout_training; do
  timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "$grp" -p no:cacheprovider > gpurun_out/m_$grp.log 2>&1
  c=$?; echo "== $grp exit $c: $(tail -1 gpurun_out/m_$grp.log)"
  if [ $c -ne 0 ]; then rc=1; grep -E "^E  |Error|error" gpurun_out/m_$grp.log | head -20; fi
done
timeout 300 python tools/prof_kernels.py > gpurun_out/kern_times.log 2>&1; cat gpurun_out/kern_times.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'])"
exit $rc
