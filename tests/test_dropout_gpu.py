"""Dropout forward / backward mask consistency, site by site (B200 only).

No dropout mask is ever stored: every kernel regenerates it from (Philox seed, element index), so a seed or index
mismatch between a forward kernel and its backward twin would train silently wrong.  Each test below INFERS the mask
the forward kernel applied (from its output, or by probing it with inputs that make the output a read-out of the
mask), then checks that the backward kernel's result equals a plain fp32 torch backward that uses exactly that mask.
Sites (engine.py `_seed(sv, layer, site)`): dense-output GEMM epilogue <-> LayerNorm backward (sites 3, 4),
band attention probabilities (site 1), global CLS row probabilities (site 2), embeddings (site 0).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from recformer_b200 import ops

DEV = "cuda"
P_DROP = 0.1
KEEP_SCALE = 1.0 / (1.0 - P_DROP)


def rnd(*shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


# ------------------------------------------------------------------ GEMM epilogue <-> LayerNorm backward
@pytest.mark.parametrize("T", [256, 1000])
def test_gemm_epilogue_dropout_mask_is_the_one_layernorm_bwd_regenerates(T):
    """forward: pre = residual + dropout(A W^T + b) (rf_gemm_bf16 epilogue, HF:1069-1070);
    backward: d(dense out) = mask/(1-p) * dLN (rf_layernorm_bwd's dx_dropped output, same seed)."""
    E, K, seed = 768, 768, 0x1234567
    A, W = rnd(T, K, seed=1), rnd(E, K, seed=2, scale=0.05)
    bias = rnd(E, seed=3, dtype=torch.float32)
    res = rnd(T, E, seed=4, dtype=torch.float32)
    pre_p = ops.gemm(A, W, bias=bias, residual=res, drop_p=P_DROP, drop_seed=seed, out_dtype=torch.float32)
    pre_0 = ops.gemm(A, W, bias=bias, residual=res, out_dtype=torch.float32)
    dense_p, dense_0 = pre_p - res, pre_0 - res
    fwd_keep = dense_p.abs() > 0.5 * dense_0.abs()                    # dropped entries are exactly `res`
    assert 0.88 < fwd_keep.float().mean().item() < 0.92
    assert relerr(dense_p[fwd_keep], dense_0[fwd_keep] * KEEP_SCALE) < 1e-3
    # backward twin
    gamma = 1 + rnd(E, seed=5, scale=0.1, dtype=torch.float32)
    beta = rnd(E, seed=6, scale=0.1, dtype=torch.float32)
    stats = torch.empty(T, 2, dtype=torch.float32, device=DEV)
    ops.layernorm_fwd(pre_p, gamma, beta, 1e-5, stats=stats)
    dy = rnd(T, E, seed=7)
    dg, db, dbias = (torch.zeros(E, dtype=torch.float32, device=DEV) for _ in range(3))
    dx = torch.empty(T, E, dtype=torch.bfloat16, device=DEV)
    dxd = torch.full((T, E), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.layernorm_bwd(dy, pre_p, stats, gamma, dg, db, dx=dx, dx_dropped=dxd, drop_p=P_DROP, drop_seed=seed, d_bias=dbias)
    live = dx.float().abs() > 0                                         # where the mask is observable
    bwd_keep = dxd.float() != 0
    assert torch.equal(bwd_keep[live], fwd_keep[live]), "LayerNorm backward regenerated a different dropout mask"
    assert relerr(dxd.float()[fwd_keep], dx.float()[fwd_keep] * KEEP_SCALE) < 1e-2
    # the dense layer's bias gradient is the column sum of the DROPPED gradient
    xf = pre_p.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xf, (E,), gamma, beta, 1e-5).backward(dy.float())
    assert relerr(dbias, (xf.grad * fwd_keep * KEEP_SCALE).sum(0)) < 2e-3
    # the epilogue can also SAVE its mask (one byte per 8 columns) for the backward to read back: same forward
    # output, bit-identical backward, and the saved bits are really used (inverted bits give another result)
    saved = torch.zeros(T, E // 8, dtype=torch.uint8, device=DEV)
    pre_m = ops.gemm(A, W, bias=bias, residual=res, drop_p=P_DROP, drop_seed=seed, out_dtype=torch.float32, drop_mask=saved)
    assert torch.equal(pre_m, pre_p)
    bits = ((saved[:, :, None] >> torch.arange(8, device=DEV, dtype=torch.uint8)) & 1).bool().view(T, E)
    assert torch.equal(bits[live], fwd_keep[live])
    dx2 = torch.empty_like(dx)
    dxd2 = torch.full_like(dxd, float("nan"))
    dbias2 = torch.zeros_like(dbias)
    ops.layernorm_bwd(dy, pre_p, stats, gamma, torch.zeros_like(dg), torch.zeros_like(db), dx=dx2, dx_dropped=dxd2,
                      drop_p=P_DROP, drop_seed=seed + 99, d_bias=dbias2, drop_mask=saved)      # the seed is not consulted
    assert torch.equal(dxd2.view(torch.int16), dxd.view(torch.int16)) and torch.equal(dx2.view(torch.int16), dx.view(torch.int16))
    assert relerr(dbias2, dbias) < 1e-5
    ops.layernorm_bwd(dy, pre_p, stats, gamma, torch.zeros_like(dg), torch.zeros_like(db), dx=dx2, dx_dropped=dxd2,
                      drop_p=P_DROP, drop_seed=seed, d_bias=dbias2, drop_mask=~saved)
    assert (dxd2.float() != 0).ne(bwd_keep)[live].float().mean().item() > 0.9
    # and a different seed gives a different mask (the comparison above is not vacuous)
    other = ops.gemm(A, W, bias=bias, residual=res, drop_p=P_DROP, drop_seed=seed + 1, out_dtype=torch.float32)
    assert ((other - res).abs() > 0.5 * dense_0.abs()).ne(fwd_keep).float().mean().item() > 0.1


# ------------------------------------------------------------------------------------- band attention
def _band_allowed(mask012, L, w):
    valid, glob = mask012 > 0, mask012 > 1
    idx = torch.arange(L, device=mask012.device)
    band = (idx[:, None] - idx[None, :]).abs() <= w
    return (band[None] & (valid & ~glob)[:, None, :]) | glob[:, None, :], valid


def _probe_band_keep(mask, allowed, B, L, H, w, seed):
    """Reads the (dropped, scaled) probability matrix A[b,h,i,j] out of rf_band_attn_fwd: with q = 0 the softmax is
    uniform over the allowed keys, and one-hot value rows v_j = e_(j mod 64) make ctx[i, d] the sum of A[i, j] over
    the allowed keys with j mod 64 = d.  One pass per residue of (j // 64) modulo the number of 64-key blocks a band can
    touch, plus one for the global CLS key, leaves at most ONE allowed key per (i, d) in every pass."""
    E = H * 64
    j = torch.arange(L, device=DEV)
    A = torch.zeros(B, H, L, L, device=DEV)
    ncls = (2 * w) // 64 + 2           # a band of 2w+1 keys touches at most this many 64-key blocks
    for cls in [((j // 64) % ncls == c) & (j != 0) for c in range(ncls)] + [j == 0]:
        qkv = torch.zeros(B, L, 3, H, 64, device=DEV)
        qkv[:, :, 2] = (torch.nn.functional.one_hot(j % 64, 64).float() * cls[:, None].float())[None, :, None, :]
        ctx, _ = ops.band_attn_fwd(qkv.view(B * L, 3 * E).to(torch.bfloat16), mask, B, L, H, w, drop_p=P_DROP, drop_seed=seed)
        out = ctx.view(B, L, H, 64).permute(0, 2, 1, 3).float()                  # (B,H,L,64)
        A += out[..., j % 64] * (allowed & cls[None, None, :])[:, None].float()   # (B,H,L,L)
    return A


@pytest.mark.parametrize("L,w", [(256, 32), (256, 64), (320, 96), (512, 128)])
def test_band_attention_backward_regenerates_the_forward_dropout_mask(L, w):
    """w = 32: one forward / one backward pass; 64 and 128: single-pass forward kernels, 65-offset backward segments;
    96: TWO forward segments whose key origin is not 8-aligned (the absolute-coordinate mask bits are re-aligned)."""
    B, H, seed = 2, 12, 0xABCDEF01
    E = H * 64
    mask = torch.ones(B, L, dtype=torch.uint8, device=DEV)
    mask[:, 0] = 2
    mask[1, L - 56:] = 0
    allowed, valid = _band_allowed(mask.long(), L, w)
    A = _probe_band_keep(mask, allowed, B, L, H, w, seed)
    n_allowed = allowed.sum(-1).clamp_min(1).float()                                            # (B, L)
    keep = A * n_allowed[:, None, :, None] / KEEP_SCALE                                          # ~1 kept / 0 dropped
    live = (allowed & valid[:, :, None])[:, None].expand(-1, H, -1, -1).clone()
    live[:, :, 0, :] = False                                                                     # global row: not the band's
    assert ((keep[live] - 1).abs() < 0.02).logical_or(keep[live] == 0).all(), "probe did not isolate single keys"
    frac = (keep[live] > 0.5).float().mean().item()
    assert 0.88 < frac < 0.92, frac
    keepb = (keep > 0.5) & live
    # real inputs, same seed: the forward output must be softmax * mask / (1-p) @ V with the PROBED mask ...
    qkv = rnd(B * L, 3 * E, seed=5)
    qkv[:, :E] *= 0.35
    ctx, lse = ops.band_attn_fwd(qkv, mask, B, L, H, w, drop_p=P_DROP, drop_seed=seed)
    qf = qkv.float().clone().requires_grad_(True)
    q, k, v = qf.view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)).masked_fill(~allowed[:, None], float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, -1), nan=0.0).masked_fill(~valid[:, None, :, None], 0.0)
    ref_ctx = ((p * keepb * KEEP_SCALE) @ v).transpose(1, 2).reshape(B, L, E)
    got = ctx.view(B, L, E).float()
    assert (got[:, 1:] - ref_ctx[:, 1:]).abs().max() < 3e-2          # P' = p/(1-p) is rounded to bf16 before PV
    # ... and the backward kernel (which regenerates the mask) must match autograd through that same mask
    dctx = rnd(B * L, E, seed=7)
    dqkv = torch.full((B * L, 3 * E), float("nan"), dtype=torch.bfloat16, device=DEV)
    scratch = torch.empty(B * L, 2 * E, dtype=torch.float32, device=DEV)
    ops.band_attn_bwd(qkv, mask, B, L, H, w, ctx, lse, dctx, dqkv, scratch, drop_p=P_DROP, drop_seed=seed)
    g = dctx.float().view(B, L, E).clone()
    g[:, 0] = 0
    ref_ctx.backward(g)
    ref = qf.grad.clone()
    ref[:, :E] *= 0.125
    for name, sl in (("dq", slice(0, E)), ("dk", slice(E, 2 * E)), ("dv", slice(2 * E, 3 * E))):
        assert relerr(dqkv[:, sl], ref[:, sl]) < 2.5e-2, name
    # a backward pass with another seed is far off: the check above discriminates
    ops.band_attn_bwd(qkv, mask, B, L, H, w, ctx, lse, dctx, dqkv, scratch, drop_p=P_DROP, drop_seed=seed + 1)
    assert relerr(dqkv[:, 2 * E:], ref[:, 2 * E:]) > 0.1


# ------------------------------------------------------------------------------------- global CLS row
def test_global_row_backward_regenerates_the_forward_dropout_mask():
    B, L, H, seed = 3, 128, 12, 0x5151
    E = H * 64
    mask = torch.ones(B, L, dtype=torch.uint8, device=DEV)
    mask[:, 0] = 2
    mask[1, 90:] = 0
    valid = mask > 0
    zeros = torch.zeros(E, E, device=DEV)
    eye = torch.eye(E, device=DEV)
    zb = torch.zeros(E, device=DEV)
    # probe: W_q = 0 -> uniform probabilities; W_v = I, x_j = one-hot(j mod 64) in every head -> ctx row 0 reads p'
    keep = torch.zeros(B, H, L, device=DEV)
    n_valid = valid.sum(-1).float()
    for half in range(L // 64):
        x = torch.zeros(B, L, H, 64, device=DEV)
        jj = torch.arange(half * 64, half * 64 + 64, device=DEV)
        x[:, jj, :, :] = torch.eye(64, device=DEV)[None, :, None, :]
        ctx = torch.zeros(B * L, E, dtype=torch.bfloat16, device=DEV)
        ops.global_attn_fwd(x.view(B * L, E).to(torch.bfloat16), mask, zeros, zb, zeros, eye.contiguous(), zb, B, L, H, ctx,
                            drop_p=P_DROP, drop_seed=seed)
        row = ctx.view(B, L, H, 64)[:, 0].float()                                               # (B,H,64) = p'[b,h,j]
        keep[:, :, jj] = row * n_valid[:, None, None] / KEEP_SCALE
    live = valid[:, None, :].expand(-1, H, -1)
    assert ((keep[live] - 1).abs() < 0.02).logical_or(keep[live] == 0).all()
    frac = (keep[live] > 0.5).float().mean().item()
    assert 0.85 < frac < 0.95, frac
    keepb = ((keep > 0.5) & live).float()
    # real weights / inputs with the same seed
    x = rnd(B * L, E, seed=1)
    Wq, Wk, Wv = (rnd(E, E, seed=s, scale=0.03, dtype=torch.float32) for s in (2, 3, 4))
    bq, bk, bv = (rnd(E, seed=s, scale=0.1, dtype=torch.float32) for s in (5, 6, 7))
    ctx = torch.zeros(B * L, E, dtype=torch.bfloat16, device=DEV)
    saved = ops.global_attn_fwd(x, mask, Wq, bq, Wk, Wv, bv, B, L, H, ctx, drop_p=P_DROP, drop_seed=seed)
    xf = x.float().view(B, L, E).clone().requires_grad_(True)
    P = {n: t.clone().requires_grad_(True) for n, t in (("Wq", Wq), ("bq", bq), ("Wk", Wk), ("Wv", Wv), ("bv", bv))}
    qg = ((xf[:, 0] @ P["Wq"].T + P["bq"]) / 8).view(B, H, 1, 64)
    kg = (xf @ P["Wk"].T + bk).view(B, L, H, 64).transpose(1, 2)
    vg = (xf @ P["Wv"].T + P["bv"]).view(B, L, H, 64).transpose(1, 2)
    s = (qg @ kg.transpose(-1, -2)).masked_fill(~valid[:, None, None, :], float("-inf"))
    pd = torch.softmax(s, -1) * keepb[:, :, None, :] * KEEP_SCALE
    out = (pd @ vg).reshape(B, E)
    assert (ctx.view(B, L, E)[:, 0].float() - out).abs().max() < 1.5e-2
    dctx = rnd(B * L, E, seed=9)
    out.backward(dctx.float().view(B, L, E)[:, 0])
    dx0 = torch.zeros(B * L, E, dtype=torch.bfloat16, device=DEV)
    dx = dx0.clone()
    grads = {n: torch.zeros_like(t) for n, t in (("Wq", Wq), ("bq", bq), ("Wk", Wk), ("Wv", Wv), ("bv", bv))}
    ops.global_attn_bwd(x, mask, Wq, bq, Wk, Wv, bv, B, L, H, dctx, saved, dx, grads["Wq"], grads["bq"], grads["Wk"],
                        grads["Wv"], grads["bv"], drop_p=P_DROP, drop_seed=seed)
    ref_dx = xf.grad.view(B * L, E)
    assert (dx.float() - ref_dx).abs().max() < 0.03 * ref_dx.abs().max() + 2e-4
    for n in ("Wq", "bq", "Wk", "Wv", "bv"):
        assert relerr(grads[n], P[n].grad) < 5e-3, n


# ------------------------------------------------------------------------------------------ embeddings
def test_embedding_backward_regenerates_the_forward_dropout_mask():
    from oracle import recformer_oracle as O
    cfg = O.OracleConfig(vocab_size=3000, num_hidden_layers=1, attention_window=[64], max_position_embeddings=600)
    sd = {k: v.to(DEV) for k, v in O.make_state_dict(cfg, seed=3).items()}
    B, L, Lp, E, seed = 4, 200, 256, 768, 0x777
    batch = {k: v.to(DEV) for k, v in O.make_batch(cfg, B, L, seed=1, ragged=True).items()}
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    pos, _ = ops.prepare_inputs(batch["input_ids"], batch["attention_mask"], batch["global_attention_mask"], Lp, 1, err)
    p = "embeddings."
    tabs = [sd[p + n + ".weight"] for n in ("word_embeddings", "position_embeddings", "token_type_embeddings",
                                            "item_position_embeddings", "LayerNorm")] + [sd[p + "LayerNorm.bias"]]
    args = (batch["input_ids"], batch["token_type_ids"], batch["item_position_ids"], pos, *tabs, Lp, 1, 1e-5)
    out0 = torch.empty(B * Lp, E, dtype=torch.float32, device=DEV)
    outp = torch.empty(B * Lp, E, dtype=torch.float32, device=DEV)
    ops.embed_ln_fwd(*args, err, out32=out0)
    ops.embed_ln_fwd(*args, err, drop_p=P_DROP, drop_seed=seed, out32=outp)
    keep = outp.abs() > 0.5 * out0.abs()
    assert 0.88 < keep.float().mean().item() < 0.92
    assert relerr(outp[keep], out0[keep] * KEEP_SCALE) < 1e-3
    # backward: d_beta = sum_t mask/(1-p) * dout, d_gamma = sum_t mask/(1-p) * dout * xhat  (rows of the real tokens
    # and of the window padding alike: the kernel covers all B*Lp rows)
    dout = rnd(B * Lp, E, seed=8)
    dout.view(B, Lp, E)[:, L:] = 0                 # window-padding rows carry no gradient (the encoder output is sliced)
    gamma, beta = tabs[4], tabs[5]
    xhat = (out0 - beta) / gamma
    gt = [torch.zeros_like(t) for t in tabs[:4]]
    dgam, dbet = torch.zeros(E, device=DEV), torch.zeros(E, device=DEV)
    ops.embed_ln_bwd(dout, batch["input_ids"], batch["token_type_ids"], batch["item_position_ids"], pos, *tabs, Lp, 1, 1e-5,
                     gt[0], gt[1], gt[2], gt[3], dgam, dbet, drop_p=P_DROP, drop_seed=seed)
    dm = dout.float() * keep * KEEP_SCALE
    assert relerr(dbet, dm.sum(0)) < 2e-3
    assert relerr(dgam, (dm * xhat).sum(0)) < 2e-3
    # the token-type table gradient (4 rows, every token contributes) against autograd through LN with the same mask
    tt = torch.nn.functional.pad(batch["token_type_ids"], (0, Lp - L), value=0).reshape(-1)
    pre = torch.zeros(B * Lp, E, device=DEV)      # any pre-LN input with the right statistics: reconstruct from tables
    ids = torch.nn.functional.pad(batch["input_ids"], (0, Lp - L), value=1).reshape(-1)
    ip = torch.nn.functional.pad(batch["item_position_ids"], (0, Lp - L), value=1).reshape(-1)
    pre = (tabs[0][ids] + tabs[1][pos.reshape(-1).long()] + tabs[2][tt] + tabs[3][ip]).requires_grad_(True)
    torch.nn.functional.layer_norm(pre, (E,), gamma, beta, 1e-5).backward(dm)
    ref_type = torch.zeros_like(tabs[2]).index_add_(0, tt, pre.grad)
    assert relerr(gt[2], ref_type) < 5e-3


@pytest.mark.parametrize("L,ragged", [(256, False), (1024, True)])
def test_band_attention_saved_keep_bits_equal_the_regenerated_mask(L, ragged):
    """attention_window 64: the forward can save each row's dropout keep bits (rf_attn_args.keepbits) and the backward
    reads them back instead of regenerating the Philox stream — both backward passes must produce the SAME gradients
    bit for bit, and a backward pass fed with inverted bits must not (the bits are really used)."""
    B, H, w, seed = 3, 12, 32, 0x1234567
    E = H * 64
    mask = torch.ones(B, L, dtype=torch.uint8, device=DEV)
    mask[:, 0] = 2
    if ragged:
        mask[1, L - 200:] = 0
        mask[2, 300:] = 0
    qkv = rnd(B * L, 3 * E, seed=11)
    qkv[:, :E] *= 0.35
    dctx = rnd(B * L, E, seed=12)
    kb = ops.band_attn_keepbits(B, L, H, DEV)
    kb.fill_(-1)
    ctx0, lse0 = ops.band_attn_fwd(qkv, mask, B, L, H, w, drop_p=P_DROP, drop_seed=seed)
    ctx1, lse1 = ops.band_attn_fwd(qkv, mask, B, L, H, w, drop_p=P_DROP, drop_seed=seed, keepbits=kb)
    # (row 0 of a sequence is the global row: the band kernel leaves it to rf_global_attn_fwd)
    assert torch.equal(ctx0.view(B, L, E)[:, 1:], ctx1.view(B, L, E)[:, 1:]) and torch.equal(lse0[:, :, 1:], lse1[:, :, 1:])
    scratch = torch.empty(B * L, 2 * E, dtype=torch.float32, device=DEV)
    g0 = torch.full((B * L, 3 * E), float("nan"), dtype=torch.bfloat16, device=DEV)
    g1 = torch.full_like(g0, float("nan"))
    ops.band_attn_bwd(qkv, mask, B, L, H, w, ctx0, lse0, dctx, g0, scratch, drop_p=P_DROP, drop_seed=seed)
    ops.band_attn_bwd(qkv, mask, B, L, H, w, ctx0, lse0, dctx, g1, scratch, drop_p=P_DROP, drop_seed=seed, keepbits=kb)
    assert torch.equal(g0.view(torch.int16), g1.view(torch.int16))
    g2 = torch.full_like(g0, float("nan"))
    ops.band_attn_bwd(qkv, mask, B, L, H, w, ctx0, lse0, dctx, g2, scratch, drop_p=P_DROP, drop_seed=seed, keepbits=~kb)
    assert relerr(g2[:, 2 * E:], g0[:, 2 * E:].float()) > 0.1
