/* recformer_b200 — C ABI of the B200-native Recformer encoder + scoring hot path.
 *
 * The reference (norahallqvistMK/RecFormer) is pure Python: it has no FFI/plugin layer, its
 * boundary is the `recformer` Python class API.  This header is the C-ABI
 * layer underneath our drop-in Python classes (`recformer_b200/models.py`): plain pointers and
 * sizes, no torch types.  Every entry point cites the reference code whose arithmetic it
 * replaces (`ref:` = reference repo path:line, `HF:` = transformers
 * models/longformer/modeling_longformer.py which holds the un-vendored encoder arithmetic).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - every buffer (inputs, outputs, workspaces) is allocated and owned by the caller; the
 *     library never allocates device memory and keeps no mutable global state apart from a
 *     mutex-guarded cache of TMA descriptors / function attributes;
 *   - kernels are enqueued on `stream` and never synchronise;
 *   - return value: 0 on success, negative RF_ERR_* otherwise; rf_last_error() returns a
 *     thread-local message.  No C++ exceptions cross the boundary;
 *   - `bf16` buffers are raw uint16 storage of bfloat16; ids are int64 as produced by
 *     torch.LongTensor (ref: recformer/tokenization.py:28-31).
 */
#ifndef RECFORMER_B200_H_
#define RECFORMER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* rf_stream_t; /* == cudaStream_t */

#define RF_OK 0
#define RF_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define RF_ERR_CUDA (-2)    /* CUDA runtime / driver error       */

const char* rf_last_error(void);
int rf_version(void); /* 102: rf_attn_args.keepbits, rf_gemm_args.drop_mask, rf_layernorm_bwd(..., drop_mask) added;
                         structs only ever grow at the end: zero-initialise them */
/* Number of kernels launched by this library in the calling process so far (bench.py reports
 * the per-step delta as `gpu_launches`). */
unsigned long long rf_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense projections: C[M,N] = epilogue(A (*) B) on tcgen05 tensor cores (bf16 in, fp32 accum).
 * Replaces every nn.Linear on the path — HF:503-505 (query/key/value), HF:1067 (attention
 * output dense), HF:1112 (intermediate dense), HF:1126 (output dense) — and their autograd
 * (dgrad / wgrad).
 *   a_mn_major = 0: A is [M,K] row-major (lda = row pitch, elements); 1: A is stored [K,M].
 *   b_mn_major = 0: B is [N,K] row-major (nn.Linear weight layout); 1: B is stored [K,N].
 * Epilogue, applied in this order on the fp32 accumulator v(row, col):
 *   v += bias[col]; if (col < scale_ncols) v *= scale;           (q /= sqrt(D), HF:513)
 *   epi == RF_EPI_GELU : C2 <- gelu_erf(v), C <- gelu_erf'(v) (saved for backward) (HF:1112-1115)
 *   epi == RF_EPI_DGELU: v *= aux[row,col]   (aux = the gelu' tensor saved by RF_EPI_GELU)
 *   dropout(drop_p, seed) on v;  v += residual[row,col];         (HF:1069-1070, 1128-1129)
 *   out_f32 ? (accumulate ? C += v : C = v) as fp32 : C = bf16(v)
 * split_k > 1 (fp32 output only) splits K over CTAs and accumulates with red.add; the caller
 * zeroes C beforehand.
 * ------------------------------------------------------------------------------------------ */
enum { RF_EPI_NONE = 0, RF_EPI_GELU = 1, RF_EPI_DGELU = 2 };

typedef struct rf_gemm_args {
  const void* A;
  const void* B;
  void* C;
  void* C2;             /* RF_EPI_GELU: activation output [M,N] bf16 (ldc) */
  const float* bias;    /* [N] or NULL */
  const void* residual; /* [M,N] (ldr) bf16, or fp32 when residual_f32 != 0; or NULL */
  const void* aux;      /* RF_EPI_DGELU: bf16 gelu'(pre-activation) [M,N] (ldaux) */
  int M, N, K;
  int lda, ldb, ldc, ldr, ldaux;
  int a_mn_major, b_mn_major;
  int epi;
  int out_f32;
  int accumulate;
  int split_k;
  float scale;
  int scale_ncols;
  float drop_p;
  uint64_t drop_seed;
  int residual_f32;
  /* Optional extra 64-deep k-block per sequence (dgrad layout, bf16 residual, xk_rows % 256 == 0):
   *   C[rows of sequence b] += A2[rows, 0:64] * B2[b*64:(b+1)*64, 0:N]
   * A2 [M,64] bf16, B2 [(M/xk_rows)*64, N] bf16, both row-major and dense; NULL = off. */
  int xk_rows;
  const void* A2;
  const void* B2;
  /* Optional, with drop_p > 0 and N % 8 == 0: the epilogue also SAVES the dropout mask it draws, one byte per 8
   * consecutive columns (bit e = column 8g + e kept), [M, N/8] row-major; rf_layernorm_bwd(drop_mask = ...) reads it
   * back instead of regenerating the Philox stream.  NULL = off. */
  void* drop_mask;
} rf_gemm_args;

int rf_gemm_bf16(const rf_gemm_args* args, rf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input preparation (ref: recformer/models.py:68-79 create_position_ids_from_input_ids,
 * :262-272 _merge_to_attention_mask, :210-260 _pad_to_window_size, :327-329 extended mask).
 * From the tokenizer's int64 [B,L] tensors builds, for the window-padded length Lp >= L:
 *   pos_ids [B,Lp] int32 = cumsum(ids != pad) * (ids != pad) + pad   (pad for l >= L)
 *   mask012 [B,Lp] uint8 = attention_mask * (global_attention_mask + 1)  (0 for l >= L)
 * attention_mask == NULL means all ones; global_attention_mask == NULL means no global token.
 * err_flag (int32, device) gets bit 0 set if any position other than 0 is marked global (the
 * kernels implement the tokenizer's CLS-only layout, ref: recformer/tokenization.py:97-99).
 * ------------------------------------------------------------------------------------------ */
int rf_prepare_inputs(const int64_t* input_ids, const int64_t* attention_mask, const int64_t* global_attention_mask,
                      int B, int L, int Lp, int padding_idx, int32_t* pos_ids, uint8_t* mask012, int* err_flag,
                      rf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Padding-aware execution.  The reference runs every layer over all B*L positions of a right-padded batch
 * (ref: recformer/tokenization.py:110-152 pads to the batch maximum, ref: recformer/models.py:274-356 computes them);
 * what it computes on padded positions never reaches a real token (keys are masked, everything else is row-wise) and
 * their gradients are exactly zero.  rf_row_tile_flags marks the 256-row tiles of the token axis that hold at least one
 * real token (L % 256 == 0) and lists the 128-row attention query tiles inside them; rf_set_row_activity makes that
 * the launch context of the CALLING THREAD: until it is cleared (tile_flags = NULL), rf_gemm_bf16 (CTA-pair kernel with
 * M == rows, or the weight-gradient layout with K == rows), rf_layernorm_fwd/bwd, rf_colsum_bf16, rf_embed_ln_fwd/bwd
 * and rf_band_attn_fwd/bwd launched with a matching row count skip the padding-only tiles (outputs of skipped rows are
 * left untouched; reductions over rows leave them out).  Results on real tokens are unchanged.
 * ------------------------------------------------------------------------------------------ */
int rf_row_tile_flags(const uint8_t* mask012, int B, int L, uint8_t* tile_flags /* [B*L/256] */,
                      int32_t* qtile_list /* [B*L/128] */, int32_t* n_qtiles /* [1] */, rf_stream_t stream);
int rf_set_row_activity(const uint8_t* tile_flags_or_null, long long rows, const int32_t* qtile_list,
                        const int32_t* n_qtiles);

/* ------------------------------------------------------------------------------------------
 * RecformerEmbeddings (ref: recformer/models.py:108-138): 4-table gather-sum, LayerNorm,
 * dropout — one kernel, one warp per token, 128-bit loads, warp-shuffle statistics.
 *   out[t,:] = dropout(LN(word[ids[t]] + pos[pid[t]] + type[tt[t]] + item[ip[t]]))   (bf16)
 * Tokens l in [L, Lp) are the window padding of _pad_to_window_size (ids = pad, position = pad,
 * item position = pad (sic, models.py:244), token type = 0).  pos_ids is the int32 [B,Lp] array
 * from rf_prepare_inputs (or caller-provided position ids).  Out-of-range ids set bit 1 of
 * err_flag and are clamped.
 * rf_embed_ln_bwd recomputes the sum/statistics, back-propagates LayerNorm and scatter-adds
 * into the fp32 gradient tables (any of which may be NULL = frozen, e.g. --fix_word_embedding,
 * ref: finetune.py:272-275) and accumulates dgamma/dbeta.
 * ------------------------------------------------------------------------------------------ */
typedef struct rf_embed_args {
  const int64_t* input_ids;          /* [B,L] */
  const int64_t* token_type_ids;     /* [B,L] or NULL (zeros) */
  const int64_t* item_position_ids;  /* [B,L] */
  const int32_t* pos_ids;            /* [B,Lp] */
  const float* word_emb;             /* [vocab,E] */
  const float* pos_emb;              /* [max_pos,E] */
  const float* type_emb;             /* [type_size,E] */
  const float* item_emb;             /* [max_item,E] */
  const float* ln_gamma;
  const float* ln_beta;
  int B, L, Lp, E;
  int vocab, max_pos, type_size, max_item;
  int padding_idx;
  float eps;
  float drop_p;
  uint64_t drop_seed;
} rf_embed_args;

int rf_embed_ln_fwd(const rf_embed_args* a, void* out_bf16, float* out_f32_or_null, int* err_flag,
                    rf_stream_t stream);
int rf_embed_ln_bwd(const rf_embed_args* a, const void* dout_bf16, float* d_word, float* d_pos, float* d_type,
                    float* d_item, float* d_gamma, float* d_beta, rf_stream_t stream);

/* Device-side assembly of the tokenizer's batch layout (SURVEY.md §8a Spec T; ref: recformer/tokenization.py:64-152
 * `encode(items, encode_item=False)` + `padding`) from pre-tokenised items stored as CSR arrays in HBM:
 * item i owns tokens item_tokens[item_offsets[i] .. item_offsets[i+1]) with token types item_types[..] (1 key,
 * 2 value); user u's history (oldest first) is user_items[user_offsets[u] .. user_offsets[u+1]).  Writes the five
 * int64 [B, L] tensors exactly as the reference tokenizer would (most recent item first, at most max_items items,
 * truncated to max_tokens, right-padded with pad_id / max_item_pos / 3 / 0 / 0) and, if out_len != NULL, each row's
 * unpadded length.  L must be >= the longest row (use max_tokens for pad_to_max). */
int rf_assemble_batch(const int64_t* item_offsets, const int32_t* item_tokens, const uint8_t* item_types,
                      const int64_t* user_offsets, const int64_t* user_items, int B, int L, int max_items,
                      int max_tokens, int bos_id, int pad_id, int max_item_pos, int64_t* out_ids,
                      int64_t* out_item_pos, int64_t* out_types, int64_t* out_mask, int64_t* out_global, int32_t* out_len,
                      rf_stream_t stream);

/* LayerNorm over the last dim (E = 768) of the fp32 residual stream [T,E]; replaces nn.LayerNorm
 * at HF:1070, HF:1129.  The residual stream (pre-LN sums and LN outputs) is kept in fp32 and only
 * rounded to bf16 where it becomes a tensor-core operand: y_bf16 is that operand copy, y_f32 the
 * residual copy (either may be NULL); stats[t] = (mean, rstd) are saved for backward.
 * bwd: dx (bf16) from dy (bf16), the fp32 pre-LN input x and stats; dgamma/dbeta accumulated
 * (fp32, +=).  If dx_dropped != NULL it also receives dropout-masked dx (mask regenerated from
 * drop_seed; the dense branch's gradient, HF:1069) while dx keeps the residual branch's.
 * d_bias (optional, +=) receives the column sums of that dense-branch gradient, i.e. the bias
 * gradient of the nn.Linear in front of the LayerNorm. */
/* out[n] += sum_t x[t,n] for a bf16 [T,N] matrix (bias gradients of the dense layers). */
int rf_colsum_bf16(const void* x_bf16, float* out, int T, int N, int ld, rf_stream_t stream);
int rf_layernorm_fwd(const float* x_f32, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                     float* stats, int T, int E, float eps, rf_stream_t stream);
int rf_layernorm_bwd(const void* dy_bf16, const float* x_f32, const float* stats, const float* gamma, void* dx_bf16,
                     void* dx_dropped_bf16, float drop_p, uint64_t drop_seed, float* d_gamma, float* d_beta,
                     float* d_bias, int T, int E, const uint8_t* drop_mask /* [T, E/8] saved by rf_gemm_bf16, or NULL =
                     regenerate from drop_seed */, rf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Longformer sliding-window attention with a global CLS token (SURVEY.md §8a Spec A;
 * HF:481-639 + helpers HF:641-961).  qkv is the fused projection output [B*L, 3*E] bf16 with
 * q already scaled by 1/sqrt(D); mask012 is the merged mask (ref: recformer/models.py:262-272)
 * as uint8 [B,L]: 0 padding, 1 local, 2 global.  Only position 0 may be global (the
 * tokenizer's layout, ref: recformer/tokenization.py:97-99).  Work unit = (batch, head,
 * 128-query tile): TMA loads, QK^T and PV on tcgen05 with TMEM accumulators, fp32 softmax;
 * attention_window 64 runs persistent kernels (one CTA per SM walks a contiguous run of tiles,
 * forward and backward), wider windows one CTA per tile (forward) / 65-offset segments (backward).
 * ctx and dqkv must be 32-byte aligned (256-bit stores).
 *   ctx  [B*L, E] bf16: attention output of every NON-global query (row 0 of each sequence
 *        is written by rf_global_attn_fwd); padded query rows are exactly zero (HF:578).
 *   lse  [B,H,L] fp32: log-sum-exp of each query row (saved for backward).
 * one_sided_window w = attention_window/2 must be a multiple of 32 and <= 256 (attention_window 64..512).
 * ------------------------------------------------------------------------------------------ */
typedef struct rf_attn_args {
  const void* qkv; /* bf16 [B*L, 3E] */
  const uint8_t* mask012;
  int B, L, H, D;
  int w; /* one-sided window */
  float drop_p;
  uint64_t drop_seed;
  void* ws; /* workspace of rf_band_attn_ws_bytes() bytes; required when w > 32, else may be NULL */
  void* keepbits; /* optional, w == 32 with drop_p > 0: [B, H, L] x 16 bytes (16-byte aligned).  rf_band_attn_fwd saves
                     the dropout keep bits of every row's window there and rf_band_attn_bwd reads them back instead of
                     regenerating the Philox stream (a third of its softmax-backward instructions).  NULL (or any other
                     window): both passes regenerate the masks from (drop_seed, row, key) as before. */
} rf_attn_args;

/* Workspace for windows wider than attention_window 64 (w > 32): the band is covered by ceil((2w+1)/65)
 * launches of the 65-key kernels on shifted keys whose partial outputs are merged through their
 * log-sum-exps (exact); the backward accumulates dQ over the same segments.  0 for w == 32. */
long long rf_band_attn_ws_bytes(int B, int L, int H, int w);

int rf_band_attn_fwd(const rf_attn_args* a, void* ctx_bf16, float* lse, rf_stream_t stream);
/* dqkv [B*L,3E] bf16 (gradient w.r.t. the UNSCALED q and k, v projections), given dctx.
 * dkv_scratch: fp32 [B*L, 2E] workspace (zeroed by the call) in which the K/V gradients of
 * overlapping tiles and of the CLS key are accumulated before being folded into dqkv. */
int rf_band_attn_bwd(const rf_attn_args* a, const void* ctx_bf16, const float* lse, const void* dctx_bf16,
                     void* dqkv_bf16, float* dkv_scratch, rf_stream_t stream);

/* Global (CLS) query row (HF:963-1056), re-associated so that key_global/value_global are
 * never applied to all L tokens (SURVEY.md §7 hard part 3):
 *   q_g = (Wqg x_cls + bqg)/sqrt(D);  u_h = Wkg[h]^T q_g[h];  s_j = u_h . x_j  (b_kg cancels in
 *   the softmax);  p = softmax_{valid j}(s);  m_h = sum_j p_j x_j;  out[h] = Wvg[h] m_h + bvg[h].
 * Writes row 0 of every sequence in ctx.  One pass over x (scores + online softmax + p'-weighted sum per 128-token
 * chunk, merged by a small kernel).  Saved for backward (all fp32): qg [B,E], u [B,H,E], p [B,H,L] = the RAW scores
 * s_hj (-inf for padded keys; the backward recomputes the probabilities from them and the log-sum-exp), mvec [B,H,E],
 * psum [2,B,H] = (sum_j dropout(p)_j, log-sum-exp of the row).  pt [B,L,16] (token-major dropout(p), 12 heads + 4 pad;
 * 16-byte aligned) is NOT written here any more: the backward fills it.  ws: rf_global_attn_fwd_ws_bytes() scratch
 * (chunk partials), 16-byte aligned. */
typedef struct rf_global_args {
  const void* x;  /* bf16 [B*L,E] layer input */
  const uint8_t* mask012;
  const float* Wqg; const float* bqg;
  const float* Wkg;
  const float* Wvg; const float* bvg;
  int B, L, H, D;
  float drop_p;
  uint64_t drop_seed;
} rf_global_args;

long long rf_global_attn_fwd_ws_bytes(int B, int L, int H);
int rf_global_attn_fwd(const rf_global_args* a, void* ctx_bf16, float* qg, float* u, float* p, float* pt, float* mvec,
                       float* psum, float* ws, rf_stream_t stream);
/* Backward of the CLS row: reads dctx row 0; accumulates (+=) fp32 dWqg,dbqg,dWkg,dWvg,dbvg (any
 * may be NULL; dbkg is identically zero) and ADDS the dense gradient the row sends to every
 * token (through s_j and m_h) into dx (bf16 [B*L,E]).  ws: rf_global_attn_bwd_ws_bytes().
 * With dx == NULL only the weight-gradient part runs (it needs dctx but not dx, so it can overlap
 * the band-attention backward on another stream); rf_global_attn_bwd_dx then adds the token
 * gradients into dx from the workspace that call left behind. */
long long rf_global_attn_bwd_ws_bytes(int B, int L, int H);
int rf_global_attn_bwd(const rf_global_args* a, const void* dctx_bf16, const float* qg, const float* u,
                       const float* p, const float* pt, const float* mvec, const float* psum, void* dx_bf16,
                       float* dWqg, float* dbqg, float* dWkg, float* dWvg, float* dbvg, float* ws, rf_stream_t stream);
/* The three *_global weight-gradient outer products of rf_global_attn_bwd as a separate launch (autograd of
 * HF:963-1056, weight part): call rf_global_attn_bwd with dWqg = dWkg = dWvg = NULL, then this with the same ws —
 * nothing on the backward's critical path waits for it. */
int rf_global_attn_bwd_wgrad(const rf_global_args* a, const float* qg, const float* mvec, const float* ws, float* dWqg,
                             float* dWkg, float* dWvg, rf_stream_t stream);
int rf_global_attn_bwd_dx(const rf_global_args* a, const float* u, const float* pt, void* dx_bf16, const float* ws,
                          rf_stream_t stream);
/* (autograd of HF:963-1056, token-gradient part.)  Alternative to rf_global_attn_bwd_dx: packs the same token gradients as the operands of the extra
 * k-block of rf_gemm_bf16 (rf_gemm_args.A2/B2), so that the QKV dgrad GEMM adds them while it
 * produces dx:  cf [B*L,64] bf16 = (p' | 0 | ds | e_cls | 0),  dmu [B*64,E] bf16 = (dm | 0 | u | dx_cls | 0). */
int rf_global_attn_bwd_xk(const rf_global_args* a, const float* u, const float* pt, const float* ws, void* cf_bf16,
                          void* dmu_bf16, rf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Scoring (SURVEY.md §8a Spec S; ref: recformer/models.py:358-369,533-545) and metrics
 * (Spec R; ref: utils.py:76-107).
 * ------------------------------------------------------------------------------------------ */
/* y[n,:] = bf16(x[n,:] / max(||x[n,:]||, 1e-8)); fp32 or bf16 input. */
int rf_normalize_rows(const void* x, int x_is_bf16, void* y_bf16, float* norms_or_null, long long N, int E,
                      rf_stream_t stream);
/* logits[b,n] = (xn[b] . yn[n]) / temp as fp32 [B,N] (xn, yn L2-normalised bf16). */
int rf_cosine_logits(const void* xn_bf16, const void* yn_bf16, float* logits, int B, long long N, int E, float temp,
                     rf_stream_t stream);
/* Fused cosine-GEMM + temperature + per-CTA top-k: the (B,N) logits never reach HBM.
 * Outputs per user the k best (score fp32 desc, id int32; ties -> lower id) over this table
 * (ids offset by id_base for a shard), plus label_score[b] = logit of labels[b] if that id
 * lies in [id_base, id_base+N), else -inf.  ws must hold rf_cosine_topk_ws_bytes(). */
long long rf_cosine_topk_ws_bytes(int B, long long N, int k);
int rf_cosine_topk(const void* xn_bf16, const void* yn_bf16, int B, long long N, int E, float temp, int k,
                   int id_base, const int64_t* labels_or_null, float* topk_scores, int32_t* topk_ids,
                   float* label_score, void* ws, rf_stream_t stream);
/* Same, written as ONE packed fp32 buffer [B, 2k+1] = k scores | k ids (int32 bit patterns) | label score per user:
 * the unit a rank contributes to the all-gather of sharded scoring (one collective, no pack kernel). */
int rf_cosine_topk_packed(const void* xn_bf16, const void* yn_bf16, int B, long long N, int E, float temp, int k,
                          int id_base, const int64_t* labels_or_null, float* packed, void* ws, rf_stream_t stream);
/* Merge of `parts` packed buffers [parts, B, 2k+1] (the all-gather result) into the global top-k. */
/* rf_cosine_topk_packed whose result is not written locally but STORED INTO EVERY RANK'S gathered buffer over NVLink peer
 * memory — the all-gather of sharded scoring (ref: finetune.py:66-96 scores one table on one GPU; SURVEY.md §8e shards it)
 * fused into the merge kernel's epilogue.  peer_gathered[r] (host array of `world` device addresses, r = 0..world-1) is
 * rank r's (world, B, 2k+1) fp32 buffer as mapped into THIS process (torch symmetric memory `buffer_ptrs`, CUDA IPC or any
 * other peer mapping); this rank's rows go to block `rank` of each.  The caller synchronises the ranks afterwards (a
 * symmetric-memory barrier on the same stream) and runs rf_topk_merge_packed on its own gathered buffer.  2k+1 <= 32. */
#define RF_MAX_PEERS 16
int rf_cosine_topk_bcast(const void* xn_bf16, const void* yn_bf16, int B, long long N, int E, float temp, int k,
                         int id_base, const int64_t* labels_or_null, const unsigned long long* peer_gathered, int world,
                         int rank, void* workspace, rf_stream_t stream);
int rf_topk_merge_packed(const float* packed, int parts, int B, int k, float* out_scores, int32_t* out_ids,
                         float* out_label_score, rf_stream_t stream);
/* Merge `parts` lists of k (score,id) per user (e.g. the all-gathered per-GPU top-k) into the
 * global top-k; label scores are max-reduced over parts. */
int rf_topk_merge(const float* scores, const int32_t* ids, const float* label_scores, int parts, int B, int k,
                  float* out_scores, int32_t* out_ids, float* out_label_score, rf_stream_t stream);
/* Full-softmax cross entropy over cosine logits and its gradient w.r.t. the pooled vector
 * (ref: recformer/models.py:587-591): loss = mean_b(lse_n logit[b,n] - logit[b,label_b]). */
int rf_cosine_ce(const void* pooled_bf16_or_f32, int pooled_is_bf16, const void* yn_bf16, const int64_t* labels,
                 int B, long long N, int E, float temp, float* loss, float* dpooled_f32, float* ws,
                 rf_stream_t stream);
long long rf_cosine_ce_ws_bytes(int B, long long N, int E);
/* Candidate scoring (ref: recformer/models.py:539-545 with `candidates`): logits[b,c] = cos(pooled_b,
 * table[cand[b,c]]) / temp as fp32 [B,C]; pooled fp32 [B,E] (un-normalised), yn = L2-normalised bf16 table [N,E],
 * cand int64 [B,C].  rf_cosine_candidates_ce is the sampled-softmax loss of :593-597 (label in column 0,
 * mean CE against target 0) with its gradient w.r.t. pooled; ws >= B*C*4 + B*E*4 + B*8 + 768 bytes. */
int rf_cosine_candidates(const float* pooled, const void* yn_bf16, const int64_t* cand, int B, int C, long long N, int E,
                         float temp, float* logits, rf_stream_t stream);
int rf_cosine_candidates_ce(const float* pooled, const void* yn_bf16, const int64_t* cand, int B, int C, long long N, int E,
                            float temp, float* loss, float* dpooled_or_null, void* ws, rf_stream_t stream);
/* Masked-LM cross entropy of the pretraining head (ref: recformer/models.py:499-510, CrossEntropyLoss with
 * ignore_index -100 over lm_head scores): logits fp32 [M, ld] (vocabulary V <= ld, ld % 8 == 0; padded columns
 * ignored), labels int64 [M]; loss (1 float) = mean over the rows with a valid label of (lse - logit[label]);
 * dlogits (bf16 [M, ld]) = (softmax - onehot) / count, zero for ignored rows / padded columns.  `count` is a
 * device scalar holding the number of valid rows. */
int rf_mlm_ce(const float* logits, const int64_t* labels, int M, int V, long long ld, const float* count, float* loss,
              void* dlogits_bf16, rf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser plumbing on the flat parameter buffer: fused AdamW (ref: optimization.py:7-34 uses
 * torch AdamW) that also refreshes the bf16 shadow weights, and a plain fp32->bf16 cast.
 * ------------------------------------------------------------------------------------------ */
int rf_cast_f32_to_bf16(const float* x, void* y_bf16, long long n, rf_stream_t stream);
int rf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16_or_null,
                  long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                  float grad_scale, rf_stream_t stream);
/* Same update (ref: optimization.py:7-34, torch.optim.AdamW) with the per-step scalars read from device memory,
 * hp_dev = {lr, 1 - beta1^t, sqrt(1 - beta2^t), grad_scale}: the form a captured CUDA graph replays (kernel
 * arguments are frozen at capture; the reference's linear-warmup scheduler changes lr every step). */
int rf_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16_or_null,
                      long long n, float beta1, float beta2, float eps, float weight_decay, const float* hp_dev,
                      rf_stream_t stream);
/* Same update reading the gradient as bf16 — the wire format of the data-parallel gradient all-reduce
 * (recformer_b200.dist.GradSync; the reference's DeepSpeed stage-2 path reduces fp16 gradients,
 * ref: lightning_pretrain.py:143).  hp_dev may be NULL (then lr / step / grad_scale arguments are used). */
int rf_adamw_step_bf16grad(float* param, const void* grad_bf16, float* exp_avg, float* exp_avg_sq, void* shadow_bf16_or_null,
                           long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                           float grad_scale, const float* hp_dev_or_null, rf_stream_t stream);
/* rf_adamw_step (hp_dev == NULL) / rf_adamw_step_dev (hp_dev != NULL) that additionally overwrites the fp32 gradient it
 * has just consumed with zeros — the `optimizer.zero_grad()` of the reference loop (ref: finetune.py:126) folded into
 * the update, so a captured step needs no separate memset of the flat gradient buffer. */
int rf_adamw_step_zero(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16_or_null,
                       long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                       float grad_scale, const float* hp_dev_or_null, rf_stream_t stream);
/* Dropout sites: ref: recformer/models.py:134 (embeddings), HF:585,1035 (attention probabilities), HF:1069,1128
 * (dense outputs); torch draws them from its generator state.  Here the masks are Philox draws keyed by
 * (drop_seed argument XOR a library-wide nonce).  The nonce is 0
 * until this call loads it from device memory (one tiny kernel per translation unit that draws masks);
 * a captured training step advances *nonce_dev before each replay to draw fresh masks. */
int rf_set_dropout_nonce(const unsigned long long* nonce_dev, rf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RECFORMER_B200_H_ */
