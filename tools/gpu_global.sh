#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "global" -p no:cacheprovider > gpurun_out/t_global.log 2>&1; echo "== global tests exit $?: $(tail -1 gpurun_out/t_global.log)"; grep -E "^E  |Error|FAILED" gpurun_out/t_global.log | head
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/m_all.log 2>&1; echo "== model tests exit $?: $(tail -1 gpurun_out/m_all.log)"; grep -E "^E  |Error|FAILED" gpurun_out/m_all.log | head
timeout 300 python tools/prof_kernels.py global_fwd global_bwd 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_global.csv python tools/prof_kernels.py global_fwd global_bwd > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/launches_global.csv | grep "rf::"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'])"
