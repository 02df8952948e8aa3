#!/bin/bash
# 2-GPU validation: scoring tests + shard timing on one GPU, multi-GPU equality checks, bench at N=1 and N=2 on the same box
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -p no:cacheprovider --tb=short -x > gpurun_out/t_score.log 2>&1; echo "== scoring tests exit $?: $(tail -1 gpurun_out/t_score.log)"; grep -E "^E  |FAILED|^ERROR" gpurun_out/t_score.log | head
RF_PROF_ITEMS=125000 timeout 300 python tools/prof_kernels.py score_topk 2>&1 | sed 's/score_topk/score_topk_125k_shard/'
timeout 300 python tools/prof_kernels.py score_topk 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29541 tools/check_multigpu.py --full > gpurun_out/check_n2.log 2>&1; echo "== check_multigpu exit $?"; grep check_multigpu gpurun_out/check_n2.log | cut -c1-1500
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "== bench n1 exit $?"
timeout 900 $TR --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "== bench n2 exit $?"; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1])
        s = d.get("secondary") or {}
        print(f, "train", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "| eval", round(s.get("value", 0)), "ms", s.get("ms_per_pass"), "e2e", round((s.get("e2e") or {}).get("value", 0)))
        print("   checks", json.dumps(d.get("checks"))[:1200]); print("   c3", json.dumps((d.get("extras") or {}).get("pretrain_c3"))[:400])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
