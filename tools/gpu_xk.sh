#!/bin/bash
# extra-k-block dgrad: tests + A/B bench
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or global" > gpurun_out/xk_tests.log 2>&1
echo "kernel tests rc=$?" >> gpurun_out/xk_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q >> gpurun_out/xk_tests.log 2>&1
echo "model tests rc=$?" >> gpurun_out/xk_tests.log
tail -5 gpurun_out/xk_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/xk_bench_fused.json 2> gpurun_out/xk_bench_fused.err
RF_DEBUG_NO_XK=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/xk_bench_plain.json 2> gpurun_out/xk_bench_plain.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/xk_bench_fused2.json 2>> gpurun_out/xk_bench_fused.err
cat gpurun_out/xk_bench_fused.json gpurun_out/xk_bench_plain.json gpurun_out/xk_bench_fused2.json | cut -c1-200
