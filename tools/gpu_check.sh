#!/bin/bash
# Pre-commit GPU check: all gpu tests (no -x: every failure is listed), smoke, default bench.
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -rA > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest gpu exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^E  |FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-3000; tail -5 gpurun_out/bench.err
