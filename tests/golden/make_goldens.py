"""Generate tests/golden/*.pt by running the UNMODIFIED reference (under oracle/ref_shim.py).

Run in the build container only (needs /root/reference):

    python tests/golden/make_goldens.py

Weights and inputs are regenerated from seeds by oracle.recformer_oracle.make_* (numpy PCG64,
platform independent), so the fixtures hold only the reference's OUTPUTS plus the case
parameters.  The GPU box re-creates the identical inputs and compares against these.
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import recformer_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def small_cfg(**kw):
    base = dict(vocab_size=2000, num_hidden_layers=2, attention_window=[64, 64], max_position_embeddings=1030)
    base.update(kw)
    return base


def build_ref_seqrec(ocfg, sd_seed, items):
    ref = ref_shim.load_reference()
    cfg = ref_shim.reference_config(ocfg)
    cfg.item_num = items.shape[0]
    m = ref.RecformerForSeqRec(cfg).eval()
    sd = O.make_state_dict(ocfg, seed=sd_seed, prefix="longformer.")
    m.load_state_dict(sd, strict=True)
    m.init_item_embedding(items)
    return m, sd


def case_forward(name, cfg_kw, B, L, ragged, N, sd_seed=0, batch_seed=0, hidden_stride=1):
    ocfg = O.OracleConfig(**cfg_kw)
    items = O.make_item_table(N, ocfg.hidden_size, seed=1)
    m, _ = build_ref_seqrec(ocfg, sd_seed, items)
    batch = O.make_batch(ocfg, B, L, seed=batch_seed, ragged=ragged)
    t = time.time()
    with torch.no_grad():
        out = m.longformer(**batch)
        logits = m(**batch)
    dt = time.time() - t
    g = {"cfg": cfg_kw, "B": B, "L": L, "ragged": ragged, "N": N, "sd_seed": sd_seed, "batch_seed": batch_seed,
         "hidden_stride": hidden_stride,
         "pooler_output": out.pooler_output.clone(),
         "last_hidden_sample": out.last_hidden_state[:, ::hidden_stride].clone(),
         "logits": logits.clone(), "ref_seconds": dt}
    print(f"{name}: ref fwd {dt:.2f}s logits std {logits.std():.3f}")
    return g


def case_train(name, cfg_kw, B, L, N, sd_seed=0, batch_seed=0, n_samples=0):
    """Loss + gradient fingerprints from the reference's own autograd (dropout disabled by
    eval(): the reference's train-mode RNG stream cannot be reproduced, SURVEY.md §7 hard part 7).
    With n_samples > 0 every gradient tensor additionally stores `n_samples` entries at seeded random
    flat positions (`sample_idx`, `sample`) plus its abs-max, for element-wise checks of the big tensors."""
    ocfg = O.OracleConfig(**cfg_kw)
    items = O.make_item_table(N, ocfg.hidden_size, seed=1)
    m, _ = build_ref_seqrec(ocfg, sd_seed, items)
    batch = O.make_batch(ocfg, B, L, seed=batch_seed, ragged=True)
    labels = torch.from_numpy(__import__("numpy").random.default_rng(7).integers(0, N, size=B))
    loss = m(**batch, labels=labels)
    loss.backward()
    grads = {}
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        gflat = p.grad.reshape(-1)
        grads[k] = {"norm": gflat.norm().item(), "head": gflat[:32].clone(), "sum": gflat.double().sum().item()}
        if gflat.numel() <= 4096:
            grads[k]["full"] = p.grad.clone()
        if n_samples > 0:
            rng = __import__("numpy").random.default_rng(len(k) * 7919 + gflat.numel())
            idx = torch.from_numpy(rng.integers(0, gflat.numel(), size=min(n_samples, gflat.numel())))
            grads[k].update(sample_idx=idx, sample=gflat[idx].clone(), absmax=gflat.abs().max().item())
    print(f"{name}: loss {loss.item():.6f}, {len(grads)} grads")
    return {"cfg": cfg_kw, "B": B, "L": L, "N": N, "sd_seed": sd_seed, "batch_seed": batch_seed,
            "labels": labels, "loss": loss.item(), "grads": grads}


def case_pretrain(name, cfg_kw, B, La, Lb, sd_seed=0, batch_seed=0):
    """RecformerForPretraining (ref: recformer/models.py:372-520): loss, in-batch contrastive logits and
    gradient fingerprints from the reference's own autograd (eval(): dropout off; `self.training` False also
    skips the dist.all_gather branch, which the oracle emulates separately)."""
    ocfg = O.OracleConfig(**cfg_kw)
    ref = ref_shim.load_reference()
    m = ref.RecformerForPretraining(ref_shim.reference_config(ocfg)).eval()
    sd = O.make_pretrain_state_dict(ocfg, seed=sd_seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    batch = O.make_pretrain_batch(ocfg, B, La, Lb, seed=batch_seed)
    out = m(**batch)
    out.loss.backward()
    grads = {}
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        gflat = p.grad.reshape(-1)
        grads[k] = {"norm": gflat.norm().item(), "head": gflat[:32].clone(), "sum": gflat.double().sum().item()}
    print(f"{name}: loss {out.loss.item():.6f}, {len(grads)} grads, correct {int(out.cl_correct_num)}")
    return {"cfg": cfg_kw, "B": B, "La": La, "Lb": Lb, "sd_seed": sd_seed, "batch_seed": batch_seed,
            "loss": out.loss.item(), "logits": out.logits.detach().clone(), "correct": int(out.cl_correct_num),
            "grads": grads, "state_keys": sorted(m.state_dict().keys())}


def case_fraud(name, cfg_kw, B, L, sd_seed=0, batch_seed=0, pos_weight=3.0):
    """RecformerForFraudDetection + FocalLoss (ref: recformer/models.py:601-713): logits, BCE(pos_weight) loss and
    gradient fingerprints from the reference's own autograd (eval(): its three dropouts off), plus FocalLoss values
    on seeded logits / targets for three (alpha, gamma, pos_weight) settings."""
    import numpy as np
    ocfg = O.OracleConfig(**cfg_kw)
    ref = ref_shim.load_reference()
    rcfg = ref_shim.reference_config(ocfg)
    rcfg.pos_weight = pos_weight
    m = ref.RecformerForFraudDetection(rcfg).eval()
    sd = O.make_fraud_state_dict(ocfg, seed=sd_seed)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    batch = O.make_batch(ocfg, B, L, seed=batch_seed, ragged=True)
    labels = torch.from_numpy(np.random.default_rng(9).integers(0, 2, size=B))
    out = m(**batch, labels=labels)
    out["loss"].backward()
    grads = {}
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        gflat = p.grad.reshape(-1)
        grads[k] = {"norm": gflat.norm().item(), "head": gflat[:32].clone(), "sum": gflat.double().sum().item()}
        if gflat.numel() <= 4096:
            grads[k]["full"] = p.grad.clone()
    tup = m(**batch, labels=labels, return_dict=False)
    assert isinstance(tup, tuple) and torch.equal(tup[1], out["logits"])
    rng = np.random.default_rng(21)
    x = torch.from_numpy(rng.standard_normal(64).astype(np.float32) * 3)
    t = torch.from_numpy(rng.integers(0, 2, size=64).astype(np.float32))
    focal = []
    for alpha, gamma, pw in ((1, 2, None), (0.6, 2, None), (None, 1.5, 2.5)):
        f = ref.FocalLoss(alpha=alpha, gamma=gamma, pos_weight=None if pw is None else torch.tensor(pw))
        focal.append({"alpha": alpha, "gamma": gamma, "pos_weight": pw, "value": f(x, t).item()})
    print(f"{name}: loss {out['loss'].item():.6f}, logits {out['logits'].tolist()}, {len(grads)} grads, focal "
          f"{[round(f['value'], 5) for f in focal]}")
    return {"cfg": cfg_kw, "B": B, "L": L, "sd_seed": sd_seed, "batch_seed": batch_seed, "pos_weight": pos_weight,
            "labels": labels, "loss": out["loss"].item(), "logits": out["logits"].detach().clone(), "grads": grads,
            "state_keys": sorted(m.state_dict().keys()), "focal_x": x, "focal_t": t, "focal": focal}


def case_ranker():
    ru = ref_shim.load_reference_utils()
    import numpy as np
    rng = np.random.default_rng(11)
    out = []
    for B, N, tie in ((64, 4000, False), (32, 500, True), (16, 100, "all")):
        s = torch.from_numpy(rng.standard_normal((B, N), dtype=np.float32))
        if tie is True:
            s = torch.round(s * 2) / 2
        if tie == "all":
            s = torch.zeros_like(s)
        labels = torch.from_numpy(rng.integers(0, N, size=(B, 1)))
        res = ru.Ranker([10, 50])(s, labels)
        out.append({"B": B, "N": N, "tie": tie, "scores": s.half() if tie is False else s, "labels": labels,
                    "metrics": res})
        if tie is False:   # keep the fixture exact: recompute on the stored (half-rounded) scores
            s2 = s.half().float()
            out[-1]["metrics"] = ru.Ranker([10, 50])(s2, labels)
    return out


def case_tokenizer():
    tk = ref_shim.load_reference_tokenization()

    class Stub:
        bos_token_id = 0
        pad_token_id = 1

    ocfg = O.OracleConfig()
    stub = Stub()
    stub.config = ocfg
    stub.encode = lambda items, encode_item=True: tk.RecformerTokenizer.encode(stub, items, encode_item)
    stub.padding = lambda item_batch, pad_to_max: tk.RecformerTokenizer.padding(stub, item_batch, pad_to_max)
    import numpy as np
    rng = np.random.default_rng(5)
    users = []
    for u in range(5):
        n_items = int(rng.integers(1, 70))
        items = []
        for _ in range(n_items):
            ln = int(rng.integers(3, 97))
            items.append([[int(x) for x in rng.integers(3, 50265, size=ln)], [int(x) for x in rng.integers(1, 3, size=ln)]])
        users.append(items)
    import copy
    enc = tk.RecformerTokenizer.batch_encode(stub, copy.deepcopy(users), encode_item=False, pad_to_max=False)
    enc_max = tk.RecformerTokenizer.batch_encode(stub, copy.deepcopy(users), encode_item=False, pad_to_max=True)
    return {"users": users, "batch": enc, "batch_pad_to_max": enc_max}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    path = os.path.join(OUT, "reference_goldens.pt")
    if "--only-pretrain" in sys.argv:      # add the pretraining case to the existing fixture file
        g = torch.load(path, weights_only=False)
        g["pretrain_small"] = case_pretrain("pretrain_small", small_cfg(), B=4, La=300, Lb=97)
        torch.save(g, path)
        print("updated", path, os.path.getsize(path) / 1e6, "MB")
        return
    if "--only-fraud" in sys.argv:         # add the classification-head case (RecformerForFraudDetection, FocalLoss)
        g = torch.load(path, weights_only=False)
        g["fraud_small"] = case_fraud("fraud_small", small_cfg(), B=4, L=300)
        torch.save(g, path)
        print("updated", path, os.path.getsize(path) / 1e6, "MB")
        return
    if "--only-train12" in sys.argv:       # add the 12-layer C2-shaped training case (BASELINE configs[1] shape)
        g = torch.load(path, weights_only=False)
        g["train_c2_12layer"] = case_train("train_c2_12layer", dict(), B=4, L=1024, N=5000, batch_seed=11, n_samples=512)
        torch.save(g, path)
        print("updated", path, os.path.getsize(path) / 1e6, "MB")
        return
    g = {}
    g["fwd_small_ragged"] = case_forward("fwd_small_ragged", small_cfg(), B=3, L=200, ragged=True, N=300, hidden_stride=3)
    g["fwd_small_dense"] = case_forward("fwd_small_dense", small_cfg(), B=2, L=256, ragged=False, N=300, hidden_stride=3)
    g["fwd_small_short"] = case_forward("fwd_small_short", small_cfg(), B=4, L=97, ragged=True, N=64, hidden_stride=2)
    for w in (128, 256, 512):
        g[f"fwd_window_{w}"] = case_forward(f"fwd_window_{w}", small_cfg(num_hidden_layers=1, attention_window=[w]),
                                            B=2, L=1024, ragged=True, N=64, hidden_stride=16)
    g["fwd_c1_full"] = case_forward("fwd_c1_full", dict(), B=8, L=1024, ragged=False, N=1000, hidden_stride=64)
    g["fwd_c1_ragged"] = case_forward("fwd_c1_ragged", dict(), B=4, L=1000, ragged=True, N=1000, hidden_stride=50,
                                      batch_seed=3)
    g["train_small"] = case_train("train_small", small_cfg(), B=3, L=200, N=50)
    g["pretrain_small"] = case_pretrain("pretrain_small", small_cfg(), B=4, La=300, Lb=97)
    g["train_c2_12layer"] = case_train("train_c2_12layer", dict(), B=4, L=1024, N=5000, batch_seed=11, n_samples=512)
    g["fraud_small"] = case_fraud("fraud_small", small_cfg(), B=4, L=300)
    g["ranker"] = case_ranker()
    g["tokenizer"] = case_tokenizer()
    path = os.path.join(OUT, "reference_goldens.pt")
    torch.save(g, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
