#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -p no:cacheprovider --tb=short -k "global" 2>&1 | tail -3 | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/global_launches.csv python tools/prof_kernels.py global_fwd global_bwd > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/global_launches.csv',errors='replace')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
kn,mv=rows[h].index('Kernel Name'),rows[h].index('Metric Value')
seen={}
for r in rows[h+1:]:
    if len(r)>mv and 'global' in r[kn]:
        seen.setdefault(r[kn].split('(')[0][-40:],[]).append(r[mv])
for k,v in seen.items(): print(k, v[-3:])
PY
