"""torch-tensor wrappers over the C ABI (include/recformer_b200.h).

PyTorch is used for device memory and streams only; every computation below is a call into
librecformer_b200.so.  All wrappers launch on the current torch CUDA stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import AttnArgs, EmbedArgs, GemmArgs, GlobalArgs, check

EPI_NONE, EPI_GELU, EPI_DGELU = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (recformer_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def launch_count() -> int:
    return int(_lib.lib().rf_launch_count())


# ---------------------------------------------------------------------------------------------
def gemm(A, B, out=None, *, bias=None, residual=None, aux=None, out2=None, a_mn_major=False, b_mn_major=False,
         epi=EPI_NONE, out_dtype=torch.bfloat16, accumulate=False, split_k=1, scale=1.0, scale_ncols=0,
         drop_p=0.0, drop_seed=0, M=None, N=None, K=None, xk=None, drop_mask=None):
    """C[M,N] = epilogue(A (*) B).  A: [M,K] (or [K,M] if a_mn_major), B: [N,K] (or [K,N] if b_mn_major).
    xk = (A2 [M,64], B2 [M/rows*64, N], rows): extra per-sequence 64-deep k-block (dgrad layout only)."""
    _req(A, torch.bfloat16, "A"), _req(B, torch.bfloat16, "B")
    if M is None:
        M, K_a = (A.shape[1], A.shape[0]) if a_mn_major else (A.shape[0], A.shape[1])
        N, K_b = (B.shape[1], B.shape[0]) if b_mn_major else (B.shape[0], B.shape[1])
        if K_a != K_b:
            raise ValueError(f"gemm: inner dimensions differ ({K_a} vs {K_b})")
        K = K_a
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    a = GemmArgs()
    a.A, a.B, a.C, a.C2 = A.data_ptr(), B.data_ptr(), out.data_ptr(), _ptr(out2)
    a.bias, a.residual, a.aux = _ptr(bias), _ptr(residual), _ptr(aux)
    a.M, a.N, a.K = M, N, K
    a.lda, a.ldb, a.ldc = A.stride(0), B.stride(0), out.stride(0)
    a.ldr = residual.stride(0) if residual is not None else 0
    a.residual_f32 = int(residual is not None and residual.dtype == torch.float32)
    a.ldaux = aux.stride(0) if aux is not None else 0
    a.a_mn_major, a.b_mn_major, a.epi = int(a_mn_major), int(b_mn_major), epi
    a.out_f32 = int(out.dtype == torch.float32)
    a.accumulate, a.split_k = int(accumulate), split_k
    a.scale, a.scale_ncols = scale, scale_ncols
    a.drop_p, a.drop_seed = drop_p, drop_seed
    if drop_mask is not None and drop_p > 0.0:      # uint8 [M, N/8]: the epilogue saves the mask it draws
        if drop_mask.dtype != torch.uint8 or drop_mask.numel() < a.M * (a.N // 8):
            raise ValueError("gemm: drop_mask must be a uint8 tensor of M * N / 8 elements")
        a.drop_mask = drop_mask.data_ptr()
    if xk is not None:
        A2, B2, rows = xk
        _req(A2, torch.bfloat16, "A2"), _req(B2, torch.bfloat16, "B2")
        if tuple(A2.shape) != (M, 64) or tuple(B2.shape) != (M // rows * 64, N):
            raise ValueError(f"gemm: extra k-block shapes {tuple(A2.shape)}, {tuple(B2.shape)} do not fit M={M} N={N}")
        a.A2, a.B2, a.xk_rows = A2.data_ptr(), B2.data_ptr(), rows
    check(_lib.lib().rf_gemm_bf16(C.byref(a), _stream()), "rf_gemm_bf16")
    return out


# ---------------------------------------------------------------------------------------------
def prepare_inputs(input_ids, attention_mask, global_attention_mask, Lp, padding_idx, err_flag):
    _req(input_ids, torch.int64, "input_ids")
    B, L = input_ids.shape
    if attention_mask is not None:
        _req(attention_mask, torch.int64, "attention_mask")
    if global_attention_mask is not None:
        _req(global_attention_mask, torch.int64, "global_attention_mask")
    pos = torch.empty(B, Lp, dtype=torch.int32, device=input_ids.device)
    mask = torch.empty(B, Lp, dtype=torch.uint8, device=input_ids.device)
    check(_lib.lib().rf_prepare_inputs(input_ids.data_ptr(), _ptr(attention_mask), _ptr(global_attention_mask), B, L,
                                       Lp, padding_idx, pos.data_ptr(), mask.data_ptr(), err_flag.data_ptr(),
                                       _stream()), "rf_prepare_inputs")
    return pos, mask


def _embed_args(input_ids, token_type_ids, item_position_ids, pos_ids, word, posw, typew, itemw, gamma, beta, Lp,
                padding_idx, eps, drop_p, drop_seed):
    B, L = input_ids.shape
    a = EmbedArgs()
    a.input_ids, a.token_type_ids = input_ids.data_ptr(), _ptr(token_type_ids)
    a.item_position_ids, a.pos_ids = item_position_ids.data_ptr(), pos_ids.data_ptr()
    a.word_emb, a.pos_emb, a.type_emb, a.item_emb = word.data_ptr(), posw.data_ptr(), typew.data_ptr(), itemw.data_ptr()
    a.ln_gamma, a.ln_beta = gamma.data_ptr(), beta.data_ptr()
    a.B, a.L, a.Lp, a.E = B, L, Lp, word.shape[1]
    a.vocab, a.max_pos, a.type_size, a.max_item = word.shape[0], posw.shape[0], typew.shape[0], itemw.shape[0]
    a.padding_idx, a.eps, a.drop_p, a.drop_seed = padding_idx, eps, drop_p, drop_seed
    return a


def row_tile_flags(mask012, B, L, flags, qtiles, n_qtiles):
    """flags[u8, B*L/256] = 256-row tile holds a real token; qtiles[i32, B*L/128] / n_qtiles[i32, 1] = active query tiles."""
    check(_lib.lib().rf_row_tile_flags(mask012.data_ptr(), B, L, flags.data_ptr(), qtiles.data_ptr(), n_qtiles.data_ptr(),
                                       _stream()), "rf_row_tile_flags")


def set_row_activity(flags=None, rows=0, qtiles=None, n_qtiles=None):
    """Launch context of this host thread: token-major kernels skip padding-only 256-row tiles; None clears it."""
    if flags is None:
        check(_lib.lib().rf_set_row_activity(None, 0, None, None), "rf_set_row_activity")
    else:
        check(_lib.lib().rf_set_row_activity(flags.data_ptr(), rows, qtiles.data_ptr(), n_qtiles.data_ptr()),
              "rf_set_row_activity")


def embed_ln_fwd(input_ids, token_type_ids, item_position_ids, pos_ids, word, posw, typew, itemw, gamma, beta, Lp,
                 padding_idx, eps, err_flag, drop_p=0.0, drop_seed=0, out=None, out32=None):
    for t, n in ((word, "word"), (posw, "pos"), (typew, "type"), (itemw, "item"), (gamma, "gamma"), (beta, "beta")):
        _req(t, torch.float32, n)
    B = input_ids.shape[0]
    a = _embed_args(input_ids, token_type_ids, item_position_ids, pos_ids, word, posw, typew, itemw, gamma, beta, Lp,
                    padding_idx, eps, drop_p, drop_seed)
    if out is None and out32 is None:
        out = torch.empty(B * Lp, word.shape[1], dtype=torch.bfloat16, device=word.device)
    check(_lib.lib().rf_embed_ln_fwd(C.byref(a), _ptr(out), _ptr(out32), _ptr(err_flag), _stream()),
          "rf_embed_ln_fwd")
    return out if out is not None else out32


def embed_ln_bwd(dout, input_ids, token_type_ids, item_position_ids, pos_ids, word, posw, typew, itemw, gamma, beta,
                 Lp, padding_idx, eps, d_word, d_pos, d_type, d_item, d_gamma, d_beta, drop_p=0.0, drop_seed=0):
    a = _embed_args(input_ids, token_type_ids, item_position_ids, pos_ids, word, posw, typew, itemw, gamma, beta, Lp,
                    padding_idx, eps, drop_p, drop_seed)
    check(_lib.lib().rf_embed_ln_bwd(C.byref(a), dout.data_ptr(), _ptr(d_word), _ptr(d_pos), _ptr(d_type),
                                     _ptr(d_item), _ptr(d_gamma), _ptr(d_beta), _stream()), "rf_embed_ln_bwd")


def layernorm_fwd(x, gamma, beta, eps, out=None, out32=None, stats=None):
    """x: fp32 [T,E] (the residual stream); out: bf16 operand copy, out32: fp32 residual copy."""
    _req(x, torch.float32, "x")
    T, E = x.shape
    if out is None and out32 is None:
        out = torch.empty(T, E, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().rf_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(out), _ptr(out32),
                                      _ptr(stats), T, E, eps, _stream()), "rf_layernorm_fwd")
    return out if out is not None else out32


def layernorm_bwd(dy, x, stats, gamma, d_gamma, d_beta, dx=None, dx_dropped=None, drop_p=0.0, drop_seed=0,
                  d_bias=None, drop_mask=None):
    _req(x, torch.float32, "x")
    T, E = x.shape
    if dx is None:
        dx = torch.empty(T, E, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().rf_layernorm_bwd(dy.data_ptr(), x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), dx.data_ptr(),
                                      _ptr(dx_dropped), drop_p, drop_seed, _ptr(d_gamma), _ptr(d_beta), _ptr(d_bias),
                                      T, E, _ptr(drop_mask), _stream()), "rf_layernorm_bwd")
    return dx


def colsum(x, out):
    T, N = x.shape
    check(_lib.lib().rf_colsum_bf16(x.data_ptr(), out.data_ptr(), T, N, x.stride(0), _stream()), "rf_colsum_bf16")
    return out


# ---------------------------------------------------------------------------------------------
def band_attn_ws(B, L, H, w, device):
    """Scratch for windows wider than attention_window 64 (None for w == 32).  The caller OWNS the buffer: the
    engine keeps one per (B, L, w) for its whole life (a captured CUDA graph bakes the pointer in, so it must
    never be freed or shared with another engine / stream); direct callers get a fresh allocation per call."""
    nbytes = int(_lib.lib().rf_band_attn_ws_bytes(B, L, H, w))
    if nbytes == 0:
        return None
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def band_attn_keepbits(B, L, H, device):
    """Buffer for the attention-probability dropout keep bits of one layer (attention_window 64 only): written by
    band_attn_fwd(keepbits=...), read back by band_attn_bwd(keepbits=...) instead of regenerating the Philox stream."""
    return torch.empty(B * H * L * 4, dtype=torch.int32, device=device)


def _attn_args(qkv, mask012, B, L, H, w, drop_p, drop_seed, ws=None, keepbits=None):
    a = AttnArgs()
    a.qkv, a.mask012 = qkv.data_ptr(), mask012.data_ptr()
    a.B, a.L, a.H, a.D, a.w = B, L, H, 64, w
    a.drop_p, a.drop_seed = drop_p, drop_seed
    if ws is None:
        ws = band_attn_ws(B, L, H, w, qkv.device)
    elif ws.numel() < int(_lib.lib().rf_band_attn_ws_bytes(B, L, H, w)):
        raise ValueError("band attention: workspace too small for this shape")
    a.ws = _ptr(ws)
    if keepbits is not None:
        if keepbits.dtype != torch.int32 or keepbits.numel() < B * H * L * 4 or keepbits.data_ptr() % 16:
            raise ValueError("band attention: keepbits must be a 16-byte aligned int32 tensor of B*H*L*4 elements")
        a.keepbits = keepbits.data_ptr() if (w == 32 and drop_p > 0.0) else None
    return a, ws


def band_attn_fwd(qkv, mask012, B, L, H, w, ctx=None, lse=None, drop_p=0.0, drop_seed=0, ws=None, keepbits=None):
    _req(qkv, torch.bfloat16, "qkv"), _req(mask012, torch.uint8, "mask012")
    if ctx is None:
        ctx = torch.empty(B * L, H * 64, dtype=torch.bfloat16, device=qkv.device)
    if lse is None:
        lse = torch.empty(B, H, L, dtype=torch.float32, device=qkv.device)
    a, ws = _attn_args(qkv, mask012, B, L, H, w, drop_p, drop_seed, ws, keepbits)
    check(_lib.lib().rf_band_attn_fwd(C.byref(a), ctx.data_ptr(), lse.data_ptr(), _stream()), "rf_band_attn_fwd")
    return ctx, lse


def band_attn_bwd(qkv, mask012, B, L, H, w, ctx, lse, dctx, dqkv, dkv_cls, drop_p=0.0, drop_seed=0, ws=None, keepbits=None):
    a, ws = _attn_args(qkv, mask012, B, L, H, w, drop_p, drop_seed, ws, keepbits)
    check(_lib.lib().rf_band_attn_bwd(C.byref(a), ctx.data_ptr(), lse.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(),
                                      dkv_cls.data_ptr(), _stream()), "rf_band_attn_bwd")
    return dqkv


def _global_args(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, drop_p, drop_seed):
    a = GlobalArgs()
    a.x, a.mask012 = x.data_ptr(), mask012.data_ptr()
    a.Wqg, a.bqg, a.Wkg, a.Wvg, a.bvg = Wqg.data_ptr(), bqg.data_ptr(), Wkg.data_ptr(), Wvg.data_ptr(), bvg.data_ptr()
    a.B, a.L, a.H, a.D = B, L, H, 64
    a.drop_p, a.drop_seed = drop_p, drop_seed
    return a


def global_attn_saved(B, L, H, device):
    """Buffers rf_global_attn_fwd saves for the backward (+ its chunk-partial scratch "ws")."""
    E = H * 64
    f32 = dict(dtype=torch.float32, device=device)
    nws = int(_lib.lib().rf_global_attn_fwd_ws_bytes(B, L, H))
    return {"qg": torch.empty(B, E, **f32), "u": torch.empty(B, H, E, **f32),
            "p": torch.empty(B, H, L, **f32),            # raw scores s_hj (the backward recomputes p from them)
            "pt": torch.empty(B, L, 16, **f32),          # token-major dropout(p): written by the backward
            "mvec": torch.empty(B, H, E, **f32),
            "psum": torch.empty(2, B, H, **f32),         # [0] = sum_j dropout(p), [1] = log-sum-exp
            "ws": torch.empty((nws + 3) // 4, **f32)}


def global_attn_fwd(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, ctx, saved=None, drop_p=0.0, drop_seed=0):
    """Writes row 0 of every sequence of ctx; returns the tensors saved for backward."""
    if saved is None:
        saved = global_attn_saved(B, L, H, x.device)
    a = _global_args(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, drop_p, drop_seed)
    check(_lib.lib().rf_global_attn_fwd(C.byref(a), ctx.data_ptr(), saved["qg"].data_ptr(), saved["u"].data_ptr(),
                                        saved["p"].data_ptr(), saved["pt"].data_ptr(), saved["mvec"].data_ptr(),
                                        saved["psum"].data_ptr(), saved["ws"].data_ptr(), _stream()), "rf_global_attn_fwd")
    return saved


# ---------------------------------------------------------------------------------------------
def normalize_rows(x, out=None, norms=None):
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("normalize_rows: fp32 or bf16 input")
    if not x.is_cuda or not x.is_contiguous():
        raise ValueError("normalize_rows: contiguous CUDA tensor required")
    N, E = x.shape
    if out is None:
        out = torch.empty(N, E, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().rf_normalize_rows(x.data_ptr(), int(x.dtype == torch.bfloat16), out.data_ptr(), _ptr(norms), N, E,
                                       _stream()), "rf_normalize_rows")
    return out


def cosine_logits(xn, yn, temp, out=None):
    _req(xn, torch.bfloat16, "xn"), _req(yn, torch.bfloat16, "yn")
    B, E = xn.shape
    N = yn.shape[0]
    if out is None:
        out = torch.empty(B, N, dtype=torch.float32, device=xn.device)
    check(_lib.lib().rf_cosine_logits(xn.data_ptr(), yn.data_ptr(), out.data_ptr(), B, N, E, temp, _stream()),
          "rf_cosine_logits")
    return out


def cosine_topk_ws(B, N, k, device):
    """Scratch of rf_cosine_topk for (B users, N items, k): allocate once and pass as `ws=` to reuse it."""
    return torch.empty(int(_lib.lib().rf_cosine_topk_ws_bytes(B, N, k)), dtype=torch.uint8, device=device)


def cosine_topk(xn, yn, temp, k=10, id_base=0, labels=None, ws=None, out=None):
    """Fused cosine scoring + top-k.  `ws` (cosine_topk_ws) and `out` = (scores [B,k] fp32, ids [B,k] int32,
    label_score [B] fp32) may be passed in to run without any allocation."""
    _req(xn, torch.bfloat16, "xn"), _req(yn, torch.bfloat16, "yn")
    B, E = xn.shape
    N = yn.shape[0]
    dev = xn.device
    nbytes = int(_lib.lib().rf_cosine_topk_ws_bytes(B, N, k))
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if out is not None:
        scores, ids, label_score = out
        _req(scores, torch.float32, "scores"), _req(ids, torch.int32, "ids"), _req(label_score, torch.float32, "label_score")
        if tuple(scores.shape) != (B, k) or tuple(ids.shape) != (B, k) or label_score.numel() != B:
            raise ValueError("cosine_topk: output buffers do not match (B, k)")
    else:
        scores = torch.empty(B, k, dtype=torch.float32, device=dev)
        ids = torch.empty(B, k, dtype=torch.int32, device=dev)
        label_score = torch.empty(B, dtype=torch.float32, device=dev)
    if labels is not None:
        _req(labels, torch.int64, "labels")
    check(_lib.lib().rf_cosine_topk(xn.data_ptr(), yn.data_ptr(), B, N, E, temp, k, id_base, _ptr(labels),
                                    scores.data_ptr(), ids.data_ptr(), label_score.data_ptr(), ws.data_ptr(),
                                    _stream()), "rf_cosine_topk")
    return scores, ids, label_score


def cosine_topk_packed(xn, yn, temp, k=10, id_base=0, labels=None, ws=None, out=None):
    """cosine_topk written as one packed fp32 [B, 2k+1] buffer (k scores | k ids as int32 bits | label score): the
    unit each rank contributes to the all-gather of sharded scoring."""
    _req(xn, torch.bfloat16, "xn"), _req(yn, torch.bfloat16, "yn")
    B, N = xn.shape[0], yn.shape[0]
    nbytes = int(_lib.lib().rf_cosine_topk_ws_bytes(B, N, k))
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xn.device)
    if out is None:
        out = torch.empty(B, 2 * k + 1, dtype=torch.float32, device=xn.device)
    if labels is not None:
        _req(labels, torch.int64, "labels")
    check(_lib.lib().rf_cosine_topk_packed(xn.data_ptr(), yn.data_ptr(), B, N, xn.shape[1], temp, k, id_base, _ptr(labels),
                                           out.data_ptr(), ws.data_ptr(), _stream()), "rf_cosine_topk_packed")
    return out


def cosine_topk_bcast(xn, yn, temp, peer_ptrs, rank, k=10, id_base=0, labels=None, ws=None):
    """cosine_topk_packed whose merged rows are stored into block `rank` of EVERY rank's gathered (world, B, 2k+1) buffer
    (peer_ptrs: their device addresses as mapped into this process, e.g. symmetric-memory buffer_ptrs)."""
    _req(xn, torch.bfloat16, "xn"), _req(yn, torch.bfloat16, "yn")
    B, N = xn.shape[0], yn.shape[0]
    nbytes = int(_lib.lib().rf_cosine_topk_ws_bytes(B, N, k))
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xn.device)
    if labels is not None:
        _req(labels, torch.int64, "labels")
    arr = (C.c_ulonglong * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    check(_lib.lib().rf_cosine_topk_bcast(xn.data_ptr(), yn.data_ptr(), B, N, xn.shape[1], temp, k, id_base, _ptr(labels),
                                          arr, len(peer_ptrs), rank, ws.data_ptr(), _stream()), "rf_cosine_topk_bcast")
    return ws


def topk_merge_packed(packed, k):
    """packed: [parts, B, 2k+1] (all-gathered cosine_topk_packed buffers) -> (scores [B,k], ids [B,k], label [B])."""
    _req(packed, torch.float32, "packed")
    parts, B, ld = packed.shape
    if ld != 2 * k + 1:
        raise ValueError("topk_merge_packed: last dimension must be 2k+1")
    dev = packed.device
    out_s = torch.empty(B, k, dtype=torch.float32, device=dev)
    out_i = torch.empty(B, k, dtype=torch.int32, device=dev)
    out_l = torch.empty(B, dtype=torch.float32, device=dev)
    check(_lib.lib().rf_topk_merge_packed(packed.data_ptr(), parts, B, k, out_s.data_ptr(), out_i.data_ptr(),
                                          out_l.data_ptr(), _stream()), "rf_topk_merge_packed")
    return out_s, out_i, out_l


def topk_merge(scores, ids, label_scores=None):
    """scores/ids: [parts, B, k]; label_scores: [parts, B] or None."""
    parts, B, k = scores.shape
    dev = scores.device
    out_s = torch.empty(B, k, dtype=torch.float32, device=dev)
    out_i = torch.empty(B, k, dtype=torch.int32, device=dev)
    out_l = torch.empty(B, dtype=torch.float32, device=dev) if label_scores is not None else None
    check(_lib.lib().rf_topk_merge(scores.data_ptr(), ids.data_ptr(), _ptr(label_scores), parts, B, k,
                                   out_s.data_ptr(), out_i.data_ptr(), _ptr(out_l), _stream()), "rf_topk_merge")
    return out_s, out_i, out_l


def cosine_ce(pooled, yn, labels, temp, want_grad=True, ws=None):
    """Full-softmax CE over cosine logits; returns (loss[1] fp32, dpooled [B,E] fp32 or None)."""
    B, E = pooled.shape
    N = yn.shape[0]
    dev = pooled.device
    nbytes = int(_lib.lib().rf_cosine_ce_ws_bytes(B, N, E)) + 1024
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dpooled = torch.empty(B, E, dtype=torch.float32, device=dev) if want_grad else None
    check(_lib.lib().rf_cosine_ce(pooled.data_ptr(), int(pooled.dtype == torch.bfloat16), yn.data_ptr(),
                                  labels.data_ptr(), B, N, E, temp, loss.data_ptr(), _ptr(dpooled), ws.data_ptr(),
                                  _stream()), "rf_cosine_ce")
    return loss, dpooled


def cosine_candidates(pooled, yn, cand, temp):
    """logits[b,c] = cos(pooled_b, table[cand[b,c]]) / temp (fp32 [B,C]) over the normalised bf16 table."""
    _req(pooled, torch.float32, "pooled"), _req(yn, torch.bfloat16, "yn"), _req(cand, torch.int64, "candidates")
    B, E = pooled.shape
    C = cand.shape[1]
    logits = torch.empty(B, C, dtype=torch.float32, device=pooled.device)
    check(_lib.lib().rf_cosine_candidates(pooled.data_ptr(), yn.data_ptr(), cand.data_ptr(), B, C, yn.shape[0], E, temp,
                                          logits.data_ptr(), _stream()), "rf_cosine_candidates")
    return logits


def cosine_candidates_ce(pooled, yn, cand, temp, want_grad=True):
    """Sampled-softmax CE (label in column 0 of cand); returns (loss[1], dpooled fp32 [B,E] or None)."""
    _req(pooled, torch.float32, "pooled"), _req(yn, torch.bfloat16, "yn"), _req(cand, torch.int64, "candidates")
    B, E = pooled.shape
    C = cand.shape[1]
    ws = torch.empty(B * C * 4 + B * E * 4 + B * 8 + 1024, dtype=torch.uint8, device=pooled.device)
    loss = torch.empty(1, dtype=torch.float32, device=pooled.device)
    dpooled = torch.empty(B, E, dtype=torch.float32, device=pooled.device) if want_grad else None
    check(_lib.lib().rf_cosine_candidates_ce(pooled.data_ptr(), yn.data_ptr(), cand.data_ptr(), B, C, yn.shape[0], E, temp,
                                             loss.data_ptr(), _ptr(dpooled), ws.data_ptr(), _stream()),
          "rf_cosine_candidates_ce")
    return loss, dpooled


def mlm_ce(logits, labels, vocab):
    """Masked-LM CE (ignore_index -100) over fp32 logits [M, ld]; returns (loss[1], dlogits bf16 [M, ld])."""
    _req(logits, torch.float32, "logits"), _req(labels, torch.int64, "labels")
    M, ld = logits.shape
    count = (labels >= 0).sum().to(torch.float32).reshape(1)
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    dlogits = torch.empty(M, ld, dtype=torch.bfloat16, device=logits.device)
    check(_lib.lib().rf_mlm_ce(logits.data_ptr(), labels.data_ptr(), M, vocab, ld, count.data_ptr(), loss.data_ptr(),
                               dlogits.data_ptr(), _stream()), "rf_mlm_ce")
    return loss, dlogits


def cast_bf16(x, out=None):
    _req(x, torch.float32, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().rf_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "rf_cast_f32_to_bf16")
    return out


def adamw_step(param, grad, exp_avg, exp_avg_sq, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(_lib.lib().rf_adamw_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                   _ptr(shadow), param.numel(), lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                                   _stream()), "rf_adamw_step")


def adamw_step_dev(param, grad, exp_avg, exp_avg_sq, shadow, beta1, beta2, eps, weight_decay, hp):
    """adamw_step with lr / bias corrections / grad_scale read from the device tensor hp (fp32 [>=4])."""
    _req(hp, torch.float32, "hp")
    check(_lib.lib().rf_adamw_step_dev(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                       _ptr(shadow), param.numel(), beta1, beta2, eps, weight_decay, hp.data_ptr(),
                                       _stream()), "rf_adamw_step_dev")


def adamw_step_bf16grad(param, grad_bf16, exp_avg, exp_avg_sq, shadow, lr, beta1, beta2, eps, weight_decay, step,
                        grad_scale=1.0, hp=None):
    """AdamW update from bf16 gradients (the wire format of dist.GradSync); hp as in adamw_step_dev or None."""
    _req(grad_bf16, torch.bfloat16, "grad")
    check(_lib.lib().rf_adamw_step_bf16grad(param.data_ptr(), grad_bf16.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                            _ptr(shadow), param.numel(), lr, beta1, beta2, eps, weight_decay, step,
                                            grad_scale, _ptr(hp), _stream()), "rf_adamw_step_bf16grad")


def adamw_step_zero(param, grad, exp_avg, exp_avg_sq, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
                    hp=None):
    """adamw_step / adamw_step_dev (hp given) that also leaves zeros in the fp32 gradient it consumed."""
    check(_lib.lib().rf_adamw_step_zero(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                        _ptr(shadow), param.numel(), lr, beta1, beta2, eps, weight_decay, step,
                                        grad_scale, _ptr(hp), _stream()), "rf_adamw_step_zero")


def set_dropout_nonce(nonce):
    """Loads the library-wide dropout nonce from a device int64 tensor (see include/recformer_b200.h)."""
    _req(nonce, torch.int64, "nonce")
    check(_lib.lib().rf_set_dropout_nonce(nonce.data_ptr(), _stream()), "rf_set_dropout_nonce")


def global_attn_bwd_ws(B, L, H, device):
    nbytes = int(_lib.lib().rf_global_attn_bwd_ws_bytes(B, L, H))
    return torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)


def global_attn_bwd(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, dctx, saved, dx, dWqg, dbqg, dWkg, dWvg, dbvg,
                    ws=None, drop_p=0.0, drop_seed=0):
    """Accumulates the *_global weight grads and, if dx is given, adds the CLS row's dense gradient into dx
    (bf16 [B*L,E]).  With dx=None call global_attn_bwd_dx afterwards (same ws); with dWqg = dWkg = dWvg = None the
    three weight-gradient outer products are left to global_attn_bwd_wgrad (same ws)."""
    nbytes = int(_lib.lib().rf_global_attn_bwd_ws_bytes(B, L, H))
    if ws is None or ws.numel() * ws.element_size() < nbytes:
        ws = global_attn_bwd_ws(B, L, H, x.device)
    a = _global_args(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, drop_p, drop_seed)
    check(_lib.lib().rf_global_attn_bwd(C.byref(a), dctx.data_ptr(), saved["qg"].data_ptr(), saved["u"].data_ptr(),
                                        saved["p"].data_ptr(), saved["pt"].data_ptr(), saved["mvec"].data_ptr(),
                                        saved["psum"].data_ptr(), _ptr(dx), _ptr(dWqg), _ptr(dbqg), _ptr(dWkg),
                                        _ptr(dWvg), _ptr(dbvg), ws.data_ptr(), _stream()), "rf_global_attn_bwd")
    return ws


def global_attn_bwd_wgrad(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, saved, ws, dWqg, dWkg, dWvg):
    """After global_attn_bwd(..., dWqg=None, dWkg=None, dWvg=None): the three *_global weight gradients (same ws)."""
    a = _global_args(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, 0.0, 0)
    check(_lib.lib().rf_global_attn_bwd_wgrad(C.byref(a), saved["qg"].data_ptr(), saved["mvec"].data_ptr(), ws.data_ptr(),
                                              dWqg.data_ptr(), dWkg.data_ptr(), dWvg.data_ptr(), _stream()),
          "rf_global_attn_bwd_wgrad")


def global_attn_bwd_xk(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, saved, ws, cf, dmu):
    """After global_attn_bwd(dx=None): packs the CLS row's token gradients as the two bf16 operands of a
    rank-64 per-sequence update, dx[b] += cf[b] @ dmu[b] (cf [B*L,64], dmu [B*64,E]), for gemm(..., xk=)."""
    a = _global_args(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, 0.0, 0)
    check(_lib.lib().rf_global_attn_bwd_xk(C.byref(a), saved["u"].data_ptr(), saved["pt"].data_ptr(), ws.data_ptr(),
                                           cf.data_ptr(), dmu.data_ptr(), _stream()), "rf_global_attn_bwd_xk")


def global_attn_bwd_dx(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, saved, dx, ws, drop_p=0.0, drop_seed=0):
    """Second half of global_attn_bwd(dx=None): adds the token gradients into dx from the workspace."""
    a = _global_args(x, mask012, Wqg, bqg, Wkg, Wvg, bvg, B, L, H, drop_p, drop_seed)
    check(_lib.lib().rf_global_attn_bwd_dx(C.byref(a), saved["u"].data_ptr(), saved["pt"].data_ptr(), dx.data_ptr(),
                                           ws.data_ptr(), _stream()), "rf_global_attn_bwd_dx")
