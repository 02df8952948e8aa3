#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -p no:cacheprovider --tb=short -k "embed or adamw" 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_model_gpu.py -q -p no:cacheprovider --tb=short -x -k "train or short or max_length or forward" 2>&1 | tail -4 | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:"embed_ln" --csv --log-file gpurun_out/q_embed_ncu.csv python tools/prof_kernels.py embed_fwd embed_bwd > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/q_embed_ncu.csv',errors='replace')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
kn,mn,mv=rows[h].index('Kernel Name'),rows[h].index('Metric Name'),rows[h].index('Metric Value')
seen={}
for r in rows[h+1:]:
    if len(r)>mv: seen[(r[kn].split('(')[0][-30:], r[mn])]=r[mv]
for k,v in seen.items(): print(k,v)
PY
