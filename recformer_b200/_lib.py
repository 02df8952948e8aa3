"""ctypes binding of librecformer_b200.so (the C ABI declared in include/recformer_b200.h).

There is NO fallback: if the shared library is missing the import of any compute entry point
raises, telling the user to build it (`python -m recformer_b200.build`)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RF_LIB_PATH: development aid (A/B runs of two builds on the same box); the product always loads the in-tree library
LIB_PATH = os.environ.get("RF_LIB_PATH") or os.path.join(HERE, "librecformer_b200.so")

c_void_p, c_int, c_float, c_ll, c_u64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_uint64


class GemmArgs(C.Structure):
    _fields_ = [("A", c_void_p), ("B", c_void_p), ("C", c_void_p), ("C2", c_void_p), ("bias", c_void_p),
                ("residual", c_void_p), ("aux", c_void_p),
                ("M", c_int), ("N", c_int), ("K", c_int),
                ("lda", c_int), ("ldb", c_int), ("ldc", c_int), ("ldr", c_int), ("ldaux", c_int),
                ("a_mn_major", c_int), ("b_mn_major", c_int), ("epi", c_int), ("out_f32", c_int),
                ("accumulate", c_int), ("split_k", c_int), ("scale", c_float), ("scale_ncols", c_int),
                ("drop_p", c_float), ("drop_seed", c_u64), ("residual_f32", c_int),
                ("xk_rows", c_int), ("A2", c_void_p), ("B2", c_void_p), ("drop_mask", c_void_p)]


class EmbedArgs(C.Structure):
    _fields_ = [("input_ids", c_void_p), ("token_type_ids", c_void_p), ("item_position_ids", c_void_p),
                ("pos_ids", c_void_p), ("word_emb", c_void_p), ("pos_emb", c_void_p), ("type_emb", c_void_p),
                ("item_emb", c_void_p), ("ln_gamma", c_void_p), ("ln_beta", c_void_p),
                ("B", c_int), ("L", c_int), ("Lp", c_int), ("E", c_int),
                ("vocab", c_int), ("max_pos", c_int), ("type_size", c_int), ("max_item", c_int),
                ("padding_idx", c_int), ("eps", c_float), ("drop_p", c_float), ("drop_seed", c_u64)]


class AttnArgs(C.Structure):
    _fields_ = [("qkv", c_void_p), ("mask012", c_void_p), ("B", c_int), ("L", c_int), ("H", c_int), ("D", c_int),
                ("w", c_int), ("drop_p", c_float), ("drop_seed", c_u64), ("ws", c_void_p), ("keepbits", c_void_p)]


class GlobalArgs(C.Structure):
    _fields_ = [("x", c_void_p), ("mask012", c_void_p), ("Wqg", c_void_p), ("bqg", c_void_p), ("Wkg", c_void_p),
                ("Wvg", c_void_p), ("bvg", c_void_p), ("B", c_int), ("L", c_int), ("H", c_int), ("D", c_int),
                ("drop_p", c_float), ("drop_seed", c_u64)]


P = C.POINTER
_SIGS = {
    "rf_last_error": (C.c_char_p, []),
    "rf_version": (c_int, []),
    "rf_launch_count": (C.c_ulonglong, []),
    "rf_gemm_bf16": (c_int, [P(GemmArgs), c_void_p]),
    "rf_prepare_inputs": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "rf_row_tile_flags": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_set_row_activity": (c_int, [c_void_p, c_ll, c_void_p, c_void_p]),
    "rf_embed_ln_fwd": (c_int, [P(EmbedArgs), c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_embed_ln_bwd": (c_int, [P(EmbedArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "rf_assemble_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_colsum_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rf_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                 c_void_p]),
    "rf_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_u64,
                                 c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rf_band_attn_ws_bytes": (c_ll, [c_int, c_int, c_int, c_int]),
    "rf_band_attn_fwd": (c_int, [P(AttnArgs), c_void_p, c_void_p, c_void_p]),
    "rf_band_attn_bwd": (c_int, [P(AttnArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_global_attn_fwd_ws_bytes": (c_ll, [c_int, c_int, c_int]),
    "rf_global_attn_fwd": (c_int, [P(GlobalArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "rf_global_attn_bwd_ws_bytes": (c_ll, [c_int, c_int, c_int]),
    "rf_global_attn_bwd": (c_int, [P(GlobalArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "rf_global_attn_bwd_dx": (c_int, [P(GlobalArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_global_attn_bwd_wgrad": (c_int, [P(GlobalArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_global_attn_bwd_xk": (c_int, [P(GlobalArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_normalize_rows": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "rf_cosine_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_void_p]),
    "rf_cosine_topk_ws_bytes": (c_ll, [c_int, c_ll, c_int]),
    "rf_cosine_topk": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_cosine_topk_packed": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "rf_cosine_topk_bcast": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_int, c_int, c_void_p, c_void_p,
                                     c_int, c_int, c_void_p, c_void_p]),
    "rf_topk_merge_packed": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_topk_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
    "rf_cosine_candidates": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_float, c_void_p, c_void_p]),
    "rf_cosine_candidates_ce": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_float, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "rf_mlm_ce": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rf_cosine_ce_ws_bytes": (c_ll, [c_int, c_ll, c_int]),
    "rf_cosine_ce": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_void_p, c_void_p,
                             c_void_p, c_void_p]),
    "rf_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "rf_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float,
                              c_float, c_float, c_int, c_float, c_void_p]),
    "rf_adamw_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float,
                                  c_float, c_void_p, c_void_p]),
    "rf_adamw_step_bf16grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float,
                                       c_float, c_float, c_int, c_float, c_void_p, c_void_p]),
    "rf_adamw_step_zero": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float,
                                   c_float, c_float, c_int, c_float, c_void_p, c_void_p]),
    "rf_set_dropout_nonce": (c_int, [c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def lib():
    """The loaded shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"recformer_b200: {LIB_PATH} is missing; build it with `python -m recformer_b200.build` "
                "(there is no CPU or PyTorch fallback for the CUDA kernels)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)     # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().rf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"recformer_b200 {what} failed ({rc}): {msg}")
