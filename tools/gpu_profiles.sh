#!/bin/bash
# Round profile set: (1) plain bench (no profiler) must exit 0, (2) ncu launch list of the same command,
# (3) ncu --set full of the dominant kernels.  Outputs land in gpurun_out/ and are summarised into profiles/
# by tools/make_profiles.py.
R=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/${R}_bench_plain.log 2> gpurun_out/${R}_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/${R}_bench_plain.err; exit 1; }
tail -1 gpurun_out/${R}_bench_plain.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extras > gpurun_out/${R}_ncu_launch.log 2>&1; echo "launch list exit $?"
CASES="gemm_qkv gemm_up_gelu gemm_down_res gemm_dgrad_dgelu gemm_wgrad_up attn_fwd attn_fwd_nodrop attn_bwd attn_bwd_regen attn_bwd_nodrop global_fwd global_bwd ln_fwd ln_bwd ln_bwd_mask colsum_3072 embed_fwd embed_bwd adamw score_topk"
timeout 600 python tools/prof_kernels.py $CASES > gpurun_out/${R}_kernel_times.log 2>&1; cat gpurun_out/${R}_kernel_times.log
RF_PROF_ITEMS=125000 timeout 300 python tools/prof_kernels.py score_topk 2>&1 | sed 's/score_topk/score_topk_125k_shard/' | tee -a gpurun_out/${R}_kernel_times.log
timeout 300 python tools/prof_kernels.py attn_fwd_w128 attn_bwd_w128 attn_fwd_w512 attn_bwd_w512 2>&1 | tee -a gpurun_out/${R}_kernel_times.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:"gemm_pair" -s 5 -c 15 -o gpurun_out/${R}_ncu_gemm python tools/prof_kernels.py gemm_qkv gemm_up_gelu gemm_down_res gemm_dgrad_dgelu gemm_wgrad_up > /dev/null 2>&1; echo "ncu gemm exit $?"
timeout 600 $NCU -k regex:"band_attn_fwd" -s 1 -c 1 -f -o gpurun_out/${R}_ncu_attn_fwd python tools/prof_kernels.py attn_fwd > /dev/null 2>&1; echo "ncu attn fwd exit $?"
timeout 600 $NCU -k regex:"band_attn_bwd" -s 1 -c 1 -f -o gpurun_out/${R}_ncu_attn_bwd python tools/prof_kernels.py attn_fwd attn_bwd > /dev/null 2>&1; echo "ncu attn bwd exit $?"
timeout 600 $NCU -k regex:"band_attn_fwd|band_attn_bwd|attn_merge" -s 6 -c 6 -o gpurun_out/${R}_ncu_attn_wide python tools/prof_kernels.py attn_fwd_w512 attn_bwd_w512 > /dev/null 2>&1; echo "ncu attn wide exit $?"
timeout 900 $NCU -k regex:"cosine_pair" -s 1 -c 1 -o gpurun_out/${R}_ncu_score python tools/prof_kernels.py score_topk > /dev/null 2>&1; echo "ncu score 1M exit $?"
RF_PROF_ITEMS=125000 timeout 600 $NCU -k regex:"cosine_pair" -s 1 -c 1 -o gpurun_out/${R}_ncu_score_shard125k python tools/prof_kernels.py score_topk > /dev/null 2>&1; echo "ncu score shard exit $?"
timeout 600 $NCU -k regex:"embed_ln|layernorm|adamw|colsum" -s 8 -c 8 -o gpurun_out/${R}_ncu_membound python tools/prof_kernels.py ln_fwd ln_bwd colsum_3072 embed_fwd embed_bwd adamw > /dev/null 2>&1; echo "ncu membound exit $?"
# The .ncu-rep files of a full set exceed what gpurun copies back (64 MiB): summarise them HERE (tools/make_profiles.py
# writes profiles/<round>_*), ship the summaries under gpurun_out/profiles_<round>/ and keep only the small captures.
python tools/make_profiles.py $R > gpurun_out/${R}_make_profiles.log 2>&1; echo "make_profiles exit $?"
mkdir -p gpurun_out/profiles_${R}; cp profiles/${R}_* gpurun_out/profiles_${R}/
find gpurun_out -name "*.ncu-rep" -size +6M -delete
du -sh gpurun_out | tail -1
