#!/bin/bash
# ncu --set full captures: scoring kernel, pair GEMMs (qkv, up_gelu), plus plain timings
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py score_topk gemm_qkv gemm_up_gelu gemm_down_res gemm_wgrad_up global_fwd global_bwd ln_fwd ln_bwd colsum_3072 > gpurun_out/kern_times.log 2>&1; cat gpurun_out/kern_times.log
RF_PROF_ITEMS=250000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cosine_mma" -s 1 -c 1 -o gpurun_out/prof_r01_score python tools/prof_kernels.py score_topk > gpurun_out/ncu_score.log 2>&1; tail -2 gpurun_out/ncu_score.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair" -s 2 -c 2 -o gpurun_out/prof_r01_gemm_pair python tools/prof_kernels.py gemm_qkv gemm_up_gelu > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
