#!/bin/bash
# Pre-commit GPU check: all gpu tests, smoke, default bench (both arms).
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest gpu exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^E  |FAILED" gpurun_out/pytest_gpu.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1200
