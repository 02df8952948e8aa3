#!/bin/bash
mkdir -p gpurun_out
rc=0
for grp in gemm band_attention_bwd; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider > gpurun_out/t_$grp.log 2>&1
  c=$?; echo "== $grp exit $c: $(tail -1 gpurun_out/t_$grp.log)"
  if [ $c -ne 0 ]; then rc=1; grep -E "^E  |Error|error|FAILED" gpurun_out/t_$grp.log | head -20; fi
done
for grp in train_step train_gradients; do
  timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "$grp" -p no:cacheprovider > gpurun_out/m_$grp.log 2>&1
  c=$?; echo "== $grp exit $c: $(tail -1 gpurun_out/m_$grp.log)"
  if [ $c -ne 0 ]; then rc=1; grep -E "^E  |Error|error" gpurun_out/m_$grp.log | head -20; fi
done
timeout 300 python tools/prof_kernels.py gemm_qkv gemm_up_gelu gemm_down_res gemm_dgrad_dgelu attn_fwd attn_fwd_nodrop attn_bwd attn_bwd_nodrop > gpurun_out/kern_times.log 2>&1; cat gpurun_out/kern_times.log
echo "-- no dkv atomics:"; RF_DEBUG_NO_DKV_ATOMICS=1 timeout 300 python tools/prof_kernels.py attn_bwd attn_bwd_nodrop 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'])"
timeout 600 python tools/prof_kernels.py attn_fwd attn_bwd > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"band_attn_bwd" -s 1 -c 1 -o gpurun_out/prof_r01_attn_bwd python tools/prof_kernels.py attn_fwd attn_bwd > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
exit $rc
