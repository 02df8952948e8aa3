"""Pin the CPU oracle against fixtures produced by the unmodified reference
(tests/golden/make_goldens.py).  CPU only."""
import copy

import numpy as np
import pytest
import torch

from oracle import recformer_oracle as O

FWD_CASES = ["fwd_small_ragged", "fwd_small_dense", "fwd_small_short", "fwd_window_128", "fwd_window_256",
             "fwd_window_512", "fwd_c1_ragged"]


@pytest.mark.parametrize("name", FWD_CASES)
def test_forward_matches_reference(goldens, name):
    g = goldens[name]
    cfg = O.OracleConfig(**g["cfg"])
    sd = O.make_state_dict(cfg, seed=g["sd_seed"], prefix="longformer.")
    batch = O.make_batch(cfg, g["B"], g["L"], seed=g["batch_seed"], ragged=g["ragged"])
    items = O.make_item_table(g["N"], cfg.hidden_size, seed=1)
    with torch.no_grad():
        h, pooled = O.model_forward(sd, cfg, prefix="longformer.", **batch)
        logits = O.similarity_score(pooled, items, cfg.temp)
    assert h.shape[1] == g["L"]
    assert (h[:, :: g["hidden_stride"]] - g["last_hidden_sample"]).abs().max() < 2e-5
    assert (pooled - g["pooler_output"]).abs().max() < 2e-5
    assert (logits - g["logits"]).abs().max() < 1e-4


def test_train_loss_and_grads_match_reference(goldens):
    g = goldens["train_small"]
    cfg = O.OracleConfig(**g["cfg"])
    sd = O.make_state_dict(cfg, seed=g["sd_seed"], prefix="longformer.")
    for k, v in sd.items():
        if v.is_floating_point():
            v.requires_grad_(True)
    batch = O.make_batch(cfg, g["B"], g["L"], seed=g["batch_seed"], ragged=True)
    items = O.make_item_table(g["N"], cfg.hidden_size, seed=1)
    loss = O.seqrec_forward(sd, cfg, batch, items, labels=g["labels"])
    assert abs(loss.item() - g["loss"]) < 1e-5
    loss.backward()
    checked = 0
    for k, ref in g["grads"].items():
        grad = sd[k].grad
        if grad is None:
            # key_global.bias: the score shift q_g.b_kg is constant over keys, softmax-invariant
            assert ref["norm"] < 1e-6, k
            continue
        tol = 1e-5 + 2e-4 * ref["norm"]
        assert abs(grad.norm().item() - ref["norm"]) < tol, k
        assert (grad.reshape(-1)[:32] - ref["head"]).abs().max() < tol, k
        if "full" in ref:
            assert (grad - ref["full"]).abs().max() < tol, k
        checked += 1
    assert checked >= 45


def test_train_12layer_c2_shape_matches_reference(goldens):
    """BASELINE configs[1] shape (12 layers, 4 ragged rows of 1024 tokens, CE over 5 000 items): the oracle's loss
    and every gradient tensor (norm, first 32 entries, 512 seeded samples) against the unmodified reference."""
    g = goldens["train_c2_12layer"]
    cfg = O.OracleConfig(**g["cfg"])
    sd = O.make_state_dict(cfg, seed=g["sd_seed"], prefix="longformer.")
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    batch = O.make_batch(cfg, g["B"], g["L"], seed=g["batch_seed"], ragged=True)
    items = O.make_item_table(g["N"], cfg.hidden_size, seed=1)
    loss = O.seqrec_forward(sd, cfg, batch, items, labels=g["labels"])
    assert abs(loss.item() - g["loss"]) < 1e-4
    loss.backward()
    checked = 0
    for k, ref in g["grads"].items():
        grad = sd[k].grad
        if grad is None:
            assert ref["norm"] < 1e-6, k
            continue
        tol = 1e-5 + 5e-4 * ref["norm"]
        assert abs(grad.norm().item() - ref["norm"]) < tol, k
        assert (grad.reshape(-1)[ref["sample_idx"]] - ref["sample"]).abs().max() < 1e-6 + 5e-4 * ref["absmax"], k
        checked += 1
    assert checked >= 240


def test_ranker_matches_reference(goldens):
    for g in goldens["ranker"]:
        got = O.ranker(g["scores"].float(), g["labels"], ks=(10, 50))
        assert np.allclose(got, g["metrics"], atol=1e-6), (g["B"], g["N"], g["tie"])
        # top-k formulation (Spec R) reproduces NDCG@10 / Recall@10 exactly
        s = g["scores"].float()
        lab = g["labels"].reshape(-1)
        top = torch.topk(s, 10, dim=-1).values
        ndcg, rec = O.topk_metrics(top, s[torch.arange(s.shape[0]), lab], 10)
        assert abs(ndcg - g["metrics"][0]) < 1e-6 and abs(rec - g["metrics"][1]) < 1e-6


def test_tokenizer_layout_matches_reference(goldens):
    g = goldens["tokenizer"]
    cfg = O.OracleConfig()
    got = O.tokenizer_batch_encode(cfg, copy.deepcopy(g["users"]), pad_to_max=False)
    assert got == g["batch"]
    got = O.tokenizer_batch_encode(cfg, copy.deepcopy(g["users"]), pad_to_max=True)
    assert got == g["batch_pad_to_max"]


def test_position_ids_and_padding_helpers():
    ids = torch.tensor([[0, 5, 6, 1, 1], [0, 7, 1, 1, 1]])
    assert O.create_position_ids_from_input_ids(ids, 1).tolist() == [[2, 3, 4, 1, 1], [2, 3, 1, 1, 1]]
    cfg = O.OracleConfig(num_hidden_layers=1, attention_window=[64])
    am = O.merge_to_attention_mask(torch.tensor([[1, 1, 1, 0, 0]]), torch.tensor([[1, 0, 0, 0, 0]]))
    assert am.tolist() == [[2, 1, 1, 0, 0]]
    pl, i2, a2, t2, p2, ip2 = O.pad_to_window_size(cfg, ids, torch.ones_like(ids), torch.zeros_like(ids), None,
                                                   torch.zeros_like(ids))
    assert pl == 59 and i2.shape[1] == 64 and int(i2[0, -1]) == 1 and int(ip2[0, -1]) == 1 and int(a2[0, -1]) == 0


def test_pretrain_step_matches_reference(goldens):
    """RecformerForPretraining (ref: recformer/models.py:372-520): loss, contrastive logits and gradient
    fingerprints of the oracle restatement vs the unmodified reference."""
    g = goldens["pretrain_small"]
    ocfg = O.OracleConfig(**g["cfg"])
    sd = O.make_pretrain_state_dict(ocfg, seed=g["sd_seed"])
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    batch = O.make_pretrain_batch(ocfg, g["B"], g["La"], g["Lb"], seed=g["batch_seed"])
    loss, cos_sim, correct = O.pretrain_forward(sd, ocfg, batch)
    assert abs(loss.item() - g["loss"]) < 1e-4
    assert (cos_sim - g["logits"]).abs().max() < 1e-3
    assert int(correct) == g["correct"]
    loss.backward()
    for k, fp in g["grads"].items():
        got = sd[k].grad
        if got is None:
            assert fp["norm"] < 1e-12, k
            continue
        assert abs(got.norm().item() - fp["norm"]) <= 2e-3 * fp["norm"] + 1e-7, k
        assert (got.reshape(-1)[:32] - fp["head"]).abs().max() <= 2e-3 * fp["head"].abs().max() + 1e-6, k
    # every parameter key of the reference module exists in the oracle's state dict
    assert set(g["state_keys"]) - set(sd) == set(), set(g["state_keys"]) - set(sd)


def test_fraud_head_and_focal_loss_match_reference(goldens):
    """RecformerForFraudDetection / FocalLoss (ref: recformer/models.py:601-713): logits, BCE(pos_weight) loss and
    gradient fingerprints of the oracle restatement vs the unmodified reference; FocalLoss of the oracle AND of the
    drop-in class (plain torch, runs on CPU) vs the reference's values."""
    g = goldens["fraud_small"]
    ocfg = O.OracleConfig(**g["cfg"])
    sd = O.make_fraud_state_dict(ocfg, seed=g["sd_seed"])
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    batch = O.make_batch(ocfg, g["B"], g["L"], seed=g["batch_seed"], ragged=True)
    loss, logits = O.fraud_forward(sd, ocfg, batch, labels=g["labels"], pos_weight=g["pos_weight"])
    assert abs(loss.item() - g["loss"]) < 1e-5
    assert (logits - g["logits"]).abs().max() < 1e-5
    loss.backward()
    checked = 0
    for k, fp in g["grads"].items():
        got = sd[k].grad
        if got is None:
            assert fp["norm"] < 1e-9, k
            continue
        tol = 1e-6 + 5e-4 * fp["norm"]
        assert abs(got.norm().item() - fp["norm"]) < tol, k
        assert (got.reshape(-1)[:32] - fp["head"]).abs().max() < tol, k
        if "full" in fp:
            assert (got - fp["full"]).abs().max() < tol, k
        checked += 1
    assert checked >= 50
    assert set(g["state_keys"]) - set(sd) == set(), set(g["state_keys"]) - set(sd)
    import recformer_b200 as rb
    for f in g["focal"]:
        pw = None if f["pos_weight"] is None else torch.tensor(f["pos_weight"])
        assert abs(O.focal_loss(g["focal_x"], g["focal_t"], f["alpha"], f["gamma"], pw).item() - f["value"]) < 1e-6
        assert abs(rb.FocalLoss(f["alpha"], f["gamma"], pw)(g["focal_x"], g["focal_t"]).item() - f["value"]) < 1e-6
    # the drop-in's parameter layout is the reference's (pos_weight is an attribute, not a buffer)
    cfg = rb.RecformerConfig(attention_window=list(ocfg.attention_window), vocab_size=ocfg.vocab_size,
                             num_hidden_layers=ocfg.num_hidden_layers, max_position_embeddings=ocfg.max_position_embeddings,
                             max_item_embeddings=ocfg.max_item_embeddings, pos_weight=g["pos_weight"])
    m = rb.RecformerForFraudDetection(cfg)
    assert sorted(m.state_dict().keys()) == g["state_keys"]
    assert float(m.pos_weight) == g["pos_weight"]
    m.load_state_dict({k: v.detach() for k, v in sd.items()}, strict=True)
