#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/m_all.log 2>&1; echo "== model tests exit $?: $(tail -1 gpurun_out/m_all.log)"; grep -E "^E  |FAILED" gpurun_out/m_all.log | head
for k in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])"; done
