#!/bin/bash
# Full GPU check: all gpu tests, smoke, default bench, ncu launch list of one bench run.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest gpu exit $?: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?: $(tail -1 gpurun_out/smoke.log)"
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "== bench ref exit $?"; tail -1 gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu.log 2>&1; echo "== ncu exit $?"
