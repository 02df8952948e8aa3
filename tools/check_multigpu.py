"""Result-equality checks of the multi-GPU paths on real hardware; run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multigpu.py [--full]

  1. sharded scoring: top-k merged over N item-id shards (rf_cosine_topk per shard -> NCCL all-gather -> rf_topk_merge)
     is bit-identical, on every rank, to the unsharded top-k of the whole table on one GPU;
  2. data-parallel training: per-layer overlapped all-reduce (dist.GradSync) of each rank's gradients, scaled by
     1/N, equals the single-GPU gradient of the concatenated global batch (dropout 0);
  3. pretraining contrastive branch (ref: recformer/models.py:475-490): NCCL all-gather of the CLS vectors with the own
     slot live; summed over ranks the gradients equal the single-process gradient of the world-sized batch.
`bench.py --gpus N` imports `run_checks` and reports the outcome in its JSON line (`checks`).  Each function returns a
dict of booleans / error magnitudes; rank 0 prints one JSON line."""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def table_chunk(chunk: int, rows: int, E: int, device, seed: int = 2):
    """Rows [chunk*rows, (chunk+1)*rows) of the synthetic N(0,1) item table, L2-normalised bf16: any rank can
    rebuild any chunk, so the sharded and the unsharded run see the same table."""
    from recformer_b200 import ops
    g = torch.Generator(device=device).manual_seed(seed * 1000 + chunk)
    return ops.normalize_rows(torch.randn(rows, E, device=device, generator=g))


def build_table(lo: int, hi: int, chunk_rows: int, E: int, device):
    assert lo % chunk_rows == 0 and hi % chunk_rows == 0
    return torch.cat([table_chunk(c, chunk_rows, E, device) for c in range(lo // chunk_rows, hi // chunk_rows)], 0)


def check_sharded_topk(device, rank, world, users=512, items=200_000, k=10):
    from recformer_b200 import dist as rdist
    from recformer_b200 import ops
    E, rows = 768, items // 40
    assert items % (rows * world) == 0, "items must split into whole chunks per rank"
    lo, hi = rank * (items // world), (rank + 1) * (items // world)
    shard = build_table(lo, hi, rows, E, device)
    xn = ops.normalize_rows(torch.randn(users, E, device=device, generator=torch.Generator(device=device).manual_seed(3)))
    labels = torch.randint(0, items, (users,), device=device, generator=torch.Generator(device=device).manual_seed(4))
    s, i, l = ops.cosine_topk(xn, shard, 0.05, k=k, id_base=lo, labels=labels)
    gs, gi, gl = rdist.all_gather_topk(s, i, l)
    ms, mi, ml = ops.topk_merge(gs, gi, gl)
    # the production path: packed per-rank result -> one all-gather -> merge of the gathered buffer in place
    packed = ops.cosine_topk_packed(xn, shard, 0.05, k=k, id_base=lo, labels=labels)
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=device)
    dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1))
    ps, pi, pl = ops.topk_merge_packed(gathered, k)
    same_packed = torch.equal(ps, ms) and torch.equal(pi, mi) and torch.equal(pl, ml)
    # the fused exchange: the scorer's merge kernel stores into every rank's symmetric buffer, barrier, merge
    peer = {"peer_exchange": "unavailable"}
    try:
        ex = rdist.PeerTopkExchange(users, k, device)
        same_peer = True
        for _ in range(3):            # three passes: both buffers, and a reuse
            ex._ws = ops.cosine_topk_bcast(xn, shard, 0.05, ex.peer_ptrs[ex._pass & 1], rank, k=k, id_base=lo, labels=labels,
                                           ws=ex._ws)
            j = ex._pass & 1
            ex._pass += 1
            ex.handles[j].barrier(channel=j)
            es, ei, el = ops.topk_merge_packed(ex.bufs[j], k)
            same_peer = same_peer and torch.equal(es, ms) and torch.equal(ei, mi) and torch.equal(el, ml)
        peer = {"peer_exchange_equals_nccl": bool(same_peer)}
    except Exception as exn:
        peer = {"peer_exchange_error": repr(exn)[:200]}
    full = build_table(0, items, rows, E, device)
    ts, ti, tl = ops.cosine_topk(xn, full, 0.05, k=k, labels=labels)
    ok = torch.tensor([int(same_packed and torch.equal(ms, ts) and torch.equal(mi, ti) and torch.equal(ml, tl)),
                       int(peer.get("peer_exchange_equals_nccl", True))], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if "peer_exchange_equals_nccl" in peer:
        peer["peer_exchange_equals_nccl"] = bool(ok[1].item())
    return {"sharded_topk_equals_unsharded": bool(ok[0].item()), "users": users, "items": items, "ranks": world, **peer}


def _small_model(device, pretraining=False, init_range=0.02):
    import recformer_b200 as rb
    cfg = rb.RecformerConfig(attention_window=[64, 64], vocab_size=1500, num_hidden_layers=2, max_position_embeddings=600,
                             max_item_embeddings=51, max_token_num=512, hidden_dropout_prob=0.0,
                             attention_probs_dropout_prob=0.0, item_num=300, initializer_range=init_range)
    torch.manual_seed(1234)                      # identical initial weights on every rank
    model = (rb.RecformerForPretraining if pretraining else rb.RecformerForSeqRec)(cfg)
    with torch.no_grad():                        # HF init leaves biases 0 / LN affine (1, 0): randomise them too
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.02)
            elif "LayerNorm.weight" in n or "layer_norm.weight" in n:
                p.normal_(1.0, 0.1)
    return model.to(device).train(), cfg


def _batch(cfg, B, L, seed, device):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, cfg.vocab_size, (B, L), generator=g)
    lens = torch.randint(L // 2, L + 1, (B,), generator=g)
    lens[0] = L
    am = (torch.arange(L)[None, :] < lens[:, None]).long()
    ids = torch.where(am.bool(), ids, torch.ones_like(ids))
    ids[:, 0] = 0
    tt = torch.where(am.bool(), torch.randint(1, 3, (B, L), generator=g), torch.full((B, L), 3))
    tt[:, 0] = 0
    ip = torch.where(am.bool(), (torch.arange(L)[None, :] // 40 + 1).expand(B, L).clamp(max=50), torch.full((B, L), 50))
    ip[:, 0] = 0
    gm = torch.zeros(B, L, dtype=torch.long)
    gm[:, 0] = 1
    b = {"input_ids": ids, "attention_mask": am, "global_attention_mask": gm, "token_type_ids": tt, "item_position_ids": ip}
    return {k: v.to(device) for k, v in b.items()}


def check_dp_gradients(device, rank, world, per_rank=2, L=256):
    from recformer_b200 import dist as rdist
    model, cfg = _small_model(device)
    model.init_item_embedding(torch.randn(cfg.item_num, 768, generator=torch.Generator().manual_seed(5)).to(device))
    full = _batch(cfg, per_rank * world, L, seed=77, device=device)
    labels = torch.randint(0, cfg.item_num, (per_rank * world,), generator=torch.Generator().manual_seed(6)).to(device)
    P = model.longformer._engine.params
    # single GPU, whole global batch (every rank computes it: same inputs, same weights)
    model(**full, labels=labels).backward()
    ref = P.grad.clone()
    P.grad.zero_()
    # data parallel: this rank's slice, per-layer overlapped all-reduce, 1/world scale (what FusedAdamW folds in)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    sync = rdist.GradSync(model)
    loss = model(**{k: v[sl] for k, v in full.items()}, labels=labels[sl])
    loss.backward()
    sync.finish()
    model.longformer._engine.grad_hook = None
    got = sync.gradient() / world       # bf16 on the wire by default: compared within bf16 rounding below
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    nerr = abs(got.norm().item() - ref.norm().item()) / ref.norm().item()
    t = torch.tensor([err, nerr], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"dp_grads_equal_single_gpu": bool(t[0].item() < 2e-2 and t[1].item() < 5e-3),
            "max_abs_err_over_absmax": float(t[0].item()), "norm_rel_err": float(t[1].item()), "ranks": world}


def check_contrastive(device, rank, world, per_rank=3):
    # N(0, 0.12) weights: at the default 0.02 the CLS vectors of a random-init encoder are collinear (cos 0.993-0.9997,
    # measured with the CPU oracle) and the contrastive gradient is a difference of nearly equal terms -- even the
    # full-batch single-GPU gradient then sits 15-25 % (of a tensor's abs-max) from the fp32 oracle; at 0.12 the
    # cosines are ~0.6 and the comparison below is well conditioned
    model, cfg = _small_model(device, pretraining=True, init_range=0.12)
    a = _batch(cfg, per_rank * world, 256, seed=91, device=device)
    b = _batch(cfg, per_rank * world, 96, seed=92, device=device)
    kw = lambda sl: dict(**{k + "_a": v[sl] for k, v in a.items()}, **{k + "_b": v[sl] for k, v in b.items()})
    P = model.longformer._engine.params
    # single process semantics on the world-sized batch: the dist branch must be off -> eval() skips it (ref :476) but
    # dropout is 0 anyway, so eval == train arithmetic here
    model.eval()
    out = model(**kw(slice(None)))
    out.loss.backward()
    ref, ref_loss = P.grad.clone(), out.loss.item()
    P.grad.zero_()
    model.train()
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    out = model(**kw(sl))
    out.loss.backward()
    dist.all_reduce(P.grad)                     # SUM over ranks of the own-slot gradients
    err = ((P.grad - ref).abs().max() / ref.abs().max()).item()
    t = torch.tensor([err, abs(out.loss.item() - ref_loss)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"contrastive_allgather_equals_world_batch": bool(t[0].item() < 3e-2 and t[1].item() < 2e-3),
            "grad_max_abs_err_over_absmax": float(t[0].item()), "loss_abs_err": float(t[1].item()),
            "logits_shape": list(out.logits.shape), "ranks": world}


def check_dp_step(device, rank, world, per_rank=2, L=256, steps=3):
    """Data-parallel optimiser step: the captured step whose AdamW updates follow each layer bucket's all-reduce on the
    aux stream (FusedAdamW.begin_overlap(sync=...)) against the plain eager sequence backward -> GradSync.finish() ->
    FusedAdamW.step(), from identical weights on identical batches; and the replicas stay bit-identical across ranks."""
    from recformer_b200 import dist as rdist
    from recformer_b200.graph import GraphedTrainStep
    from recformer_b200.optim import FusedAdamW
    out = []
    batches = None
    for mode in ("eager", "graph"):
        model, cfg = _small_model(device)
        model.longformer.strict_checks = False
        model.init_item_embedding(torch.randn(cfg.item_num, 768, generator=torch.Generator().manual_seed(5)).to(device))
        if batches is None:
            batches = []
            for s in range(steps + 1):
                full = _batch(cfg, per_rank * world, L, seed=300 + s, device=device)
                labels = torch.randint(0, cfg.item_num, (per_rank * world,), generator=torch.Generator().manual_seed(400 + s)).to(device)
                sl = slice(rank * per_rank, (rank + 1) * per_rank)
                b = {k: v[sl].contiguous() for k, v in full.items()}
                b["labels"] = labels[sl].contiguous()
                batches.append(b)
        opt = FusedAdamW(model, lr=1e-3, weight_decay=0.01)
        sync = rdist.GradSync(model, bucket_layers=1)

        def eager(b):
            loss = model(**b)
            opt.zero_grad()
            loss.backward()
            opt.step(grad_scale=1.0 / world, wait_other=sync.finish(defer_tail=True))

        eager(batches[0])
        if mode == "eager":
            for b in batches[1:]:
                eager(b)
        else:
            step = GraphedTrainStep(model, opt, batches[0], grad_scale=1.0 / world, sync=sync)
            assert step.dp_overlap
            for b in batches[1:]:
                step(b)
            torch.cuda.synchronize()
            del step
        torch.cuda.synchronize()
        model.longformer._engine.grad_hook = None
        out.append(model.longformer._engine.params.flat.clone())
    diff = (out[0] - out[1]).abs()
    # replicas: every rank must hold exactly the same parameters
    mine = out[1].clone()
    dist.broadcast(mine, src=0)
    same = torch.tensor([int(torch.equal(mine, out[1]))], device=device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    t = torch.tensor([diff.mean().item(), (diff > 1e-4).float().mean().item()], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # tolerance as in tests/test_model_gpu.py::test_graphed_train_step_matches_eager_steps (atomics in the gradient sums
    # + Adam's sign-like first steps: a wrong update rule moves every weight by ~lr)
    return {"dp_overlapped_step_equals_plain": bool(t[0].item() < 2e-5 and t[1].item() < 0.02),
            "replicas_bit_identical": bool(same.item()), "mean_abs_diff": float(t[0].item()),
            "frac_diff_gt_1e-4": float(t[1].item()), "ranks": world}


def run_checks(device, rank, world, full=False):
    res = {}
    for name, fn in (("topk", lambda: check_sharded_topk(device, rank, world)),
                     ("dp", lambda: check_dp_gradients(device, rank, world)),
                     ("dp_step", lambda: check_dp_step(device, rank, world)),
                     ("contrastive", lambda: check_contrastive(device, rank, world))):
        try:
            res[name] = fn()
        except Exception as ex:          # a failing check is reported, not fatal for the caller's own output
            res[name] = {"error": repr(ex)[:300]}
    if full:
        try:
            res["topk_full"] = check_sharded_topk(device, rank, world, users=4096, items=1_000_000)
        except Exception as ex:
            res["topk_full"] = {"error": repr(ex)[:300]}
    return res


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    res = run_checks(device, rank, world, full="--full" in sys.argv)
    if rank == 0:
        print(json.dumps({"check_multigpu": res}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ok = all(all(v for k, v in r.items() if isinstance(v, bool)) and "error" not in r for r in res.values())
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
