#!/bin/bash
# End-of-round evidence run: all gpu tests, smoke, default bench, per-kernel CUDA-event timings, the captured step's CUPTI
# timeline and fresh ncu captures of the two band attention kernels (summarised on the box: see gpu_profiles.sh).
R=${1:-r02}
cd /root/repo; mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | cut -c1-600
CASES="gemm_qkv gemm_up_gelu gemm_down_res gemm_dgrad_dgelu gemm_wgrad_up attn_fwd attn_fwd_nodrop attn_bwd attn_bwd_regen attn_bwd_nodrop global_fwd global_bwd ln_fwd ln_bwd ln_bwd_mask colsum_3072 embed_fwd embed_bwd adamw score_topk"
timeout 600 python tools/prof_kernels.py $CASES > gpurun_out/${R}_kernel_times.log 2>&1
RF_PROF_ITEMS=125000 timeout 300 python tools/prof_kernels.py score_topk 2>&1 | sed 's/score_topk/score_topk_125k_shard/' >> gpurun_out/${R}_kernel_times.log
timeout 300 python tools/prof_kernels.py attn_fwd_w128 attn_bwd_w128 attn_fwd_w512 attn_bwd_w512 >> gpurun_out/${R}_kernel_times.log 2>&1
cat gpurun_out/${R}_kernel_times.log
timeout 300 python tools/step_timeline.py 2 --graph > gpurun_out/${R}_step_timeline_graph.txt 2>&1; head -4 gpurun_out/${R}_step_timeline_graph.txt | tail -2
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:"band_attn_fwd" -s 1 -c 1 -o gpurun_out/${R}_ncu_attn_fwd python tools/prof_kernels.py attn_fwd > /dev/null 2>&1; echo "ncu attn fwd exit $?"
timeout 600 $NCU -k regex:"band_attn_bwd" -s 1 -c 1 -o gpurun_out/${R}_ncu_attn_bwd python tools/prof_kernels.py attn_fwd attn_bwd > /dev/null 2>&1; echo "ncu attn bwd exit $?"
python tools/make_profiles.py $R > gpurun_out/${R}_make_profiles.log 2>&1; echo "make_profiles exit $?"
mkdir -p gpurun_out/profiles_${R}; cp profiles/${R}_ncu_attn_fwd.txt profiles/${R}_ncu_attn_bwd.txt profiles/${R}_kernel_times.txt profiles/${R}_sass_counts.txt gpurun_out/profiles_${R}/
find gpurun_out -name "*.ncu-rep" -size +6M -delete
