"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: the packed top-k
all-gather + merge of sharded scoring, and the flat-gradient all-reduce of data-parallel training.
The per-shard scores come from the CPU oracle (Spec S); the CUDA kernels themselves are covered by
the `-m gpu` tests, here only the exchange / merge / bookkeeping runs."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"worker exited with {p.exitcode}"
    return [ret[r] for r in range(world)]


def _sharded_topk_job(rank, world):
    from oracle import recformer_oracle as O
    from recformer_b200 import dist as rdist
    from recformer_b200.metrics import Ranker, TopKRanker
    N, B, E, k = 4001, 37, 768, 10          # N not divisible by the world size: uneven shards
    items = O.make_item_table(N, E, seed=1)
    users = torch.randn(B, E, generator=torch.Generator().manual_seed(3))
    labels = torch.randint(0, N, (B,), generator=torch.Generator().manual_seed(4))
    full = O.similarity_score(users, items, 0.05)                     # (B, N) oracle logits
    full[:, 7] = full[:, 1900]                                    # a cross-shard tie: lower id must win
    lo, hi = rdist.shard_bounds(N, world, rank)
    local = full[:, lo:hi]
    s, i = torch.topk(local, k, dim=1)
    i = (i + lo).to(torch.int32)
    own = (labels >= lo) & (labels < hi)
    lab = torch.where(own, full[torch.arange(B), labels], torch.full((B,), float("-inf")))
    gs, gi, gl = rdist.all_gather_topk(s.contiguous(), i.contiguous(), lab)
    ms, mi, ml = rdist.merge_topk_reference(gs, gi, gl, k)
    # reference: unsharded top-k with ties towards the lower id, and the reference Ranker
    order = torch.argsort(full, dim=1, descending=True, stable=True)[:, :k]
    ref_s = torch.gather(full, 1, order)
    assert torch.equal(ms, ref_s)
    assert torch.equal(mi.to(torch.int64), order)
    assert torch.equal(ml, full[torch.arange(B), labels])
    dense = Ranker([10])(full, labels[:, None])
    fused = TopKRanker([10])(ms, ml)
    assert abs(dense[0] - fused[0]) < 1e-6 and abs(dense[1] - fused[1]) < 1e-6      # NDCG@10, Recall@10
    return mi.tolist()


def test_sharded_topk_allgather_merge_world2():
    out = _run(_sharded_topk_job)
    assert out[0] == out[1]           # identical result on every rank


def _shard_bounds_job(rank, world):
    from recformer_b200 import dist as rdist
    covered = []
    for r in range(world):
        covered += list(range(*rdist.shard_bounds(1_000_003, world, r)))[:: 100_000]
    lo, hi = rdist.shard_bounds(1_000_003, world, rank)
    t = torch.tensor([hi - lo])
    dist.all_reduce(t)
    return int(t.item())


def test_shard_bounds_cover_table_world2():
    assert _run(_shard_bounds_job) == [1_000_003, 1_000_003]


def _grad_allreduce_job(rank, world):
    """allreduce_gradients averages the engine's flat fp32 gradient buffer across ranks."""
    import recformer_b200 as rb
    from recformer_b200 import dist as rdist
    cfg = rb.RecformerConfig(attention_window=[64], vocab_size=120, num_hidden_layers=1, max_position_embeddings=80)
    model = rb.RecformerForSeqRec(cfg)
    P = model.longformer._engine.params
    P.grad = torch.full((P.n_total,), float(rank + 1))             # stand-in for a backward pass on this rank
    rdist.allreduce_gradients(model)
    return float(P.grad[0]), float(P.grad[-1]), P.n_total


def test_flat_gradient_allreduce_world2():
    out = _run(_grad_allreduce_job)
    assert out[0][:2] == (1.5, 1.5) and out[1][:2] == (1.5, 1.5)


def _grad_sync_job(rank, world):
    """GradSync: per-layer async all-reduces + finish() cover the flat gradient buffer exactly once."""
    import recformer_b200 as rb
    from recformer_b200 import dist as rdist
    cfg = rb.RecformerConfig(attention_window=[64, 64, 64], vocab_size=120, num_hidden_layers=3,
                             max_position_embeddings=80)
    model = rb.RecformerForSeqRec(cfg)
    eng = model.longformer._engine
    P = eng.params
    P.grad = torch.arange(P.n_total, dtype=torch.float32) % 7 + rank      # rank-dependent stand-in gradients
    expect = (torch.arange(P.n_total, dtype=torch.float32) % 7) * world + sum(range(world))
    sync = rdist.GradSync(model, bucket_layers=1)
    assert eng.grad_hook is not None
    for layer in reversed(range(3)):           # what engine.backward() does as each layer finishes
        eng.grad_hook(layer)
    assert len(sync._covered) == 6
    tail = sync.finish(defer_tail=True)            # embedding tables' all-reduce still in flight
    assert tail is not None
    tail()
    return bool(torch.equal(P.grad, expect)), len(sync._works)


def test_overlapped_grad_sync_covers_buffer_once_world2():
    assert _run(_grad_sync_job) == [(True, 0), (True, 0)]


def _grad_sync_bucket_job(rank, world):
    """Bucketed GradSync (2 layers per bucket on a 5-layer model: buckets {4}, {2,3}, {0,1}) still covers every
    element exactly once, with 2 collectives per bucket."""
    import recformer_b200 as rb
    from recformer_b200 import dist as rdist
    cfg = rb.RecformerConfig(attention_window=[64] * 5, vocab_size=120, num_hidden_layers=5, max_position_embeddings=80)
    model = rb.RecformerForSeqRec(cfg)
    eng = model.longformer._engine
    P = eng.params
    P.grad = torch.arange(P.n_total, dtype=torch.float32) % 5 + 2 * rank
    expect = (torch.arange(P.n_total, dtype=torch.float32) % 5) * world + 2 * sum(range(world))
    sync = rdist.GradSync(model, bucket_layers=2)
    counts = []
    for layer in reversed(range(5)):
        eng.grad_hook(layer)
        counts.append(len(sync._covered))
    sync.finish()
    return bool(torch.equal(P.grad, expect)), counts


def test_bucketed_grad_sync_world2():
    out = _run(_grad_sync_bucket_job)
    assert out[0] == (True, [2, 2, 4, 4, 6]) and out[1] == out[0]


def _contrastive_job(rank, world):
    """Distributed in-batch contrastive branch (ref: recformer/models.py:475-497): packed all-gather of both towers'
    CLS vectors, this rank's slot live, CE over the (B*W)^2 logits.  Checked against the oracle's single-process
    emulation (`pretrain_forward(gathered_z=...)` semantics restated on the pooled vectors): loss, logits, and the
    gradient — non-zero only through this rank's own slot, and equal to the oracle's autograd through that slot."""
    from oracle import recformer_oracle as O
    from recformer_b200.models import contrastive_head, gather_cls_with_local_grad
    B, E, temp = 5, 768, 0.05
    g = torch.Generator().manual_seed(11)
    z1_all = torch.randn(world * B, E, generator=g) + 0.5            # every rank can rebuild all ranks' vectors
    z2_all = z1_all * 0.7 + 0.7 * torch.randn(world * B, E, generator=g)
    z1 = z1_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    z2 = z2_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    a1, a2 = gather_cls_with_local_grad(z1, z2)
    assert a1.shape == (world * B, E) and torch.equal(a1.detach(), z1_all) and torch.equal(a2.detach(), z2_all)
    loss, cos_sim, correct = contrastive_head(a1, a2, temp)
    loss.backward()
    # oracle: live slot inside the gathered (constant) tensors
    r1 = z1_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    r2 = z2_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    o1 = torch.cat([z1_all[: rank * B], r1, z1_all[(rank + 1) * B:]], 0)
    o2 = torch.cat([z2_all[: rank * B], r2, z2_all[(rank + 1) * B:]], 0)
    ref_sim = O.similarity(o1.unsqueeze(1), o2.unsqueeze(0), temp)
    ref_loss = torch.nn.functional.cross_entropy(ref_sim, torch.arange(world * B))
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    assert (cos_sim - ref_sim).abs().max() < 1e-4
    assert int(correct) == int((ref_sim.argmax(1) == torch.arange(world * B)).sum())
    assert (z1.grad - r1.grad).abs().max() < 1e-6 and (z2.grad - r2.grad).abs().max() < 1e-6
    assert z1.grad.abs().max() > 0 and z2.grad.abs().max() > 0
    # summed over ranks (what the DP gradient all-reduce does downstream) the own-slot gradients reproduce the
    # gradient of the world-sized batch computed in one process
    full1, full2 = z1_all.clone().requires_grad_(True), z2_all.clone().requires_grad_(True)
    torch.nn.functional.cross_entropy(O.similarity(full1.unsqueeze(1), full2.unsqueeze(0), temp),
                                      torch.arange(world * B)).backward()
    assert (z1.grad - full1.grad[rank * B:(rank + 1) * B]).abs().max() < 1e-6
    assert (z2.grad - full2.grad[rank * B:(rank + 1) * B]).abs().max() < 1e-6
    return round(loss.item(), 5)


def test_contrastive_allgather_own_slot_gradient_world2():
    out = _run(_contrastive_job)
    assert out[0] == out[1]          # the loss is the same on every rank (same gathered batch)
