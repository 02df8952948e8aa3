#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider --tb=short -k "topk or candidate or cosine or score" 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --tb=short -k "recall" 2>&1 | tail -2 | cut -c1-300
for ni in 31250 125000 250000 1000000; do
RF_PROF_ITEMS=$ni timeout 300 python tools/prof_kernels.py score_topk 2>&1 | tail -1 | sed "s/score_topk/items_$ni/"
done
RF_PROF_ITEMS=125000 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/score_launches.csv python tools/prof_kernels.py score_topk > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/score_launches.csv',errors='replace')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
kn,mn,mv=rows[h].index('Kernel Name'),rows[h].index('Metric Name'),rows[h].index('Metric Value')
for r in rows[h+1:][-8:]:
    if len(r)>mv: print(r[kn].split('(')[0][-50:], r[mn], r[mv])
PY
