// Banded (sliding-window) attention forward with a global CLS key column, sm_100a.
//
// attention_window 64 (W = 32) runs the PERSISTENT kernel further down (band_attn_fwd_persist_kernel); the wider
// windows the one-shot kernel described here: one CTA = one (batch, head, 128-query tile), templated on the one-sided
// window W in {64, 128} (attention_window 128 / 256 in ONE pass; attention_window 512 = two W=128 segments merged
// through their log-sum-exps).  Keys j in [i0-W, i0+127+W] (NK = 128+2W rows) plus the sequence's first 16 rows
// (row 0 = the global CLS key) are TMA-loaded once (128B-swizzled, out-of-range rows zero-filled by TMA), then
//   S[128 x (NK+16)] = Q K^T          tcgen05.mma, fp32 accumulator in TMEM
//   softmax over the band + CLS column in fp32.  TWO threads per query row (warps w and w+4 share TMEM lane
//   quadrant w): each reads its half of the row's 2W+32 window columns from TMEM ONCE into registers
//   (tcgen05.ld x16), masks them with a per-row bit mask (key-valid bits of the tile AND the band position,
//   built with shifts: no per-element shared-memory flag loads), and the two halves exchange max / sum through
//   shared memory;  P (bf16) -> shared memory in the K-major UMMA layout (aliasing the dead Q/K tiles)
//   O[128 x 64] = P V                 tcgen05.mma (V is the MN-major B operand, no transpose)
//   O / rowsum -> bf16 context (32 head dims per thread); log-sum-exp saved for the backward pass.
// Dropout masks are keyed on ABSOLUTE coordinates (row, key >> 3), so the backward kernel regenerates them
// whatever its own tiling / window segmentation is.
// Semantics: SURVEY.md §8a Spec A / HF:481-639.  Global keys are removed from the band and
// re-enter through the extra column (HF:523,558-568); padded query rows produce zeros (HF:578);
// the global query row (position 0 when mask012==2) is left to rf_global_attn_fwd (HF:963-1056).
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

RF_DEFINE_NONCE_LOADER(attn_fwd)

namespace rf {

constexpr int ATT_THREADS = 256;
constexpr int HEAD_DIM = 64;

template <int W>
struct AttnFwdCfg {
  static constexpr int NK = 128 + 2 * W;       // band key rows
  static constexpr int NT = NK + 16;           // + global chunk
  static constexpr int PCH = (NT + 63) / 64;   // 64-key P chunks
  static constexpr int WIN_CH = 2 * W / 32 + 1;   // 32-column chunks a row quadrant's band can touch
  static constexpr int UNITS = WIN_CH;            // 16-column units per thread (two threads per row)
  static constexpr uint32_t Q_BYTES = 128 * 128;
  static constexpr uint32_t KV_BYTES = NT * 128;
  static constexpr uint32_t P_BYTES = PCH * 16384;
  static constexpr uint32_t REGION_A = (Q_BYTES + KV_BYTES > P_BYTES) ? (Q_BYTES + KV_BYTES) : P_BYTES;
  static constexpr uint32_t OFF_V = REGION_A;
  static constexpr uint32_t OFF_BITS = OFF_V + KV_BYTES;            // key-valid bit words
  static constexpr uint32_t OFF_RED = OFF_BITS + 64;                // [2][2][128] floats: max / sum exchange
  static constexpr uint32_t OFF_BAR = OFF_RED + 4 * 128 * 4;
  static constexpr uint32_t TOTAL = OFF_BAR + 64 + 1024;
  static constexpr uint32_t TMEM_COLS = (NT <= 256) ? 256 : 512;
  static_assert(NT <= 512, "window too large for the single-shot kernel");
  static_assert(NT <= 14 * 32, "key-valid bit words");
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

struct AttnFwdParams {
  const uint8_t* mask012;
  __nv_bfloat16* ctx;
  float* lse;
  int B, L, H;
  // window segment (windows wider than the kernel's 2W+1 keys are covered by several launches whose outputs are
  // merged through their log-sum-exps): keys are shifted by `shift` rows, the top `hi_cut` offsets of the band are
  // cut, the CLS column is only part of segment 0, and an empty row reports lse = -inf instead of 0
  int shift, hi_cut, use_cls, lse_neg_inf;
  float drop_scale;
  uint32_t drop_thresh;
  uint64_t drop_seed;
  const uint8_t* row_active;   // rf_set_row_activity: one flag per 256 token rows (L % 256 == 0), or null
};

template <int W>
__global__ void __launch_bounds__(ATT_THREADS)
band_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
                     const AttnFwdParams p) {
  using C = AttnFwdCfg<W>;
  constexpr int NK = C::NK, NT = C::NT, UNITS = C::UNITS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + C::Q_BYTES;
  uint8_t* sP = smem;  // aliases Q/K once S has been computed
  uint8_t* sV = smem + C::OFF_V;
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + C::OFF_BITS);
  float* s_red = reinterpret_cast<float*>(smem + C::OFF_RED);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int tiles_per_seq = (p.L + 127) / 128;
  const int tile = blockIdx.x % tiles_per_seq;
  const int h = (blockIdx.x / tiles_per_seq) % p.H;
  const int b = blockIdx.x / (tiles_per_seq * p.H);
  const int i0 = tile * 128;
  // a query tile inside a padding-only 256-row tile: nobody reads its output rows (before any barrier / TMEM set-up)
  if (p.row_active != nullptr && p.row_active[(static_cast<size_t>(b) * p.L + i0) >> 8] == 0) return;
  const int E = p.H * HEAD_DIM;
  const uint8_t* mrow = p.mask012 + static_cast<size_t>(b) * p.L;
  const int key0 = i0 - W + p.shift;             // absolute key index of tile column 0

  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
    // the operand loads are issued first: the TMEM allocation and the key-bit set-up below (global loads
    // of the mask) then run under their latency
    mbar_arrive_expect_tx(bar_load, C::Q_BYTES + 2 * C::KV_BYTES);
#pragma unroll
    for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 8192, &tm64, bar_load, h * HEAD_DIM, i0 + c * 64, b);
#pragma unroll
    for (int c = 0; c < NK / 64; ++c) {
      tma_load_3d(sK + c * 8192, &tm64, bar_load, E + h * HEAD_DIM, key0 + c * 64, b);
      tma_load_3d(sV + c * 8192, &tm64, bar_load, 2 * E + h * HEAD_DIM, key0 + c * 64, b);
    }
    tma_load_3d(sK + NK * 128, &tm16, bar_load, E + h * HEAD_DIM, 0, b);
    tma_load_3d(sV + NK * 128, &tm16, bar_load, 2 * E + h * HEAD_DIM, 0, b);
  }
  // the row's mask bytes are loaded now, next to the TMA loads (first used after the S MMA)
  const int r = quad * 32 + lane;       // query row of the tile == TMEM lane
  const int i = i0 + r;
  const uint8_t m_row = mrow[i < p.L ? i : 0], m_cls = mrow[0];
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  // key-valid bits of the NK band columns (bit c: key is in range, not padding, not global)
#pragma unroll 1
  for (int cb = warp * 32; cb < NK; cb += ATT_THREADS) {
    const int j = key0 + cb + lane;
    const bool inr = j >= 0 && j < p.L;
    const bool f = inr && (mrow[inr ? j : 0] == 1);
    const uint32_t word = __ballot_sync(0xffffffffu, f);
    if (lane == 0) kbits[cb >> 5] = word;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int band_hi = 2 * W - p.hi_cut;   // last in-band column offset of a row

  if (tid == 0) {
    // ---- S = Q K^T ----
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
#pragma unroll
    for (int n0 = 0; n0 < NT; n0 += 256) {
      const int n = (NT - n0) < 256 ? (NT - n0) : 256;
      const uint32_t idesc = umma_idesc_bf16(128, n, false, false);
#pragma unroll
      for (int k = 0; k < HEAD_DIM / 16; ++k)
        umma_bf16(tmem + n0, umma_smem_desc(aq + k * 32, 16, 1024), umma_smem_desc(ak + n0 * 128 + k * 32, 16, 1024),
                  idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(bar_mma);
  }
  // per-row band masks of this thread's 16-column units (independent of S: computed while the MMA runs)
  const bool row_valid = (i < p.L) && (m_row != 0);
  const int ubase = quad * 2 + part * UNITS;       // first 16-column unit of this thread (tile column / 16)
  uint32_t wmask[UNITS];
#pragma unroll
  for (int u = 0; u < UNITS; ++u) {
    const int c0 = (ubase + u) * 16;
    const uint32_t kw = (kbits[c0 >> 5] >> (c0 & 31)) & 0xFFFFu;
    const int lo = r - c0, hi = r + band_hi - c0;
    const uint32_t mlo = lo <= 0 ? 0xFFFFu : (lo >= 16 ? 0u : ((0xFFFFu << lo) & 0xFFFFu));
    const uint32_t mhi = hi >= 15 ? 0xFFFFu : (hi < 0 ? 0u : (0xFFFFu >> (15 - hi)));
    wmask[u] = row_valid ? (kw & mlo & mhi) : 0u;
  }
  const bool g_ok = p.use_cls && (m_cls == 2);
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  // ---- softmax: threads (quad, part 0) and (quad, part 1) share query row i0 + r (TMEM lane r) ----
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const float LOG2E = 1.4426950408889634f;
  uint32_t sv[UNITS][16];
#pragma unroll
  for (int u = 0; u < UNITS; ++u) tmem_ld16(lane_base + (ubase + u) * 16, sv[u]);
  uint32_t g16[16];
  if (part == 1) tmem_ld16(lane_base + NK, g16);   // warp-uniform
  tmem_ld_wait();
  float m = -INFINITY;
#pragma unroll
  for (int u = 0; u < UNITS; ++u)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float s = ((wmask[u] >> j) & 1u) ? __uint_as_float(sv[u][j]) : -INFINITY;
      sv[u][j] = __float_as_uint(s);
      m = fmaxf(m, s);
    }
  float sg = -INFINITY;
  if (part == 1 && g_ok && row_valid) sg = __uint_as_float(g16[0]);
  m = fmaxf(m, sg);
  s_red[part * 128 + r] = m;
  __syncthreads();
  m = fmaxf(s_red[r], s_red[128 + r]);
  if (m == -INFINITY) m = 0.0f;   // fully masked row: every p below is exp2(-inf) = 0
  const float m2 = m * LOG2E;

  const uint64_t rowid = (static_cast<uint64_t>(b) * p.H + h) * p.L + (i < p.L ? i : 0);
  const uint64_t rowbase = rowid * attn_drop_groups(p.L);
  float l = 0.0f;
#pragma unroll
  for (int u = 0; u < UNITS; ++u) {
    float pr[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float e = exp2f(__uint_as_float(sv[u][j]) * LOG2E - m2);
      l += e;
      pr[j] = e;
    }
    if (p.drop_thresh != 0 && wmask[u] != 0) {
      const uint32_t keep = attn_keep16(p.drop_seed, rowbase, key0 + (ubase + u) * 16, p.drop_thresh);
#pragma unroll
      for (int j = 0; j < 16; ++j) pr[j] = ((keep >> j) & 1u) ? pr[j] * p.drop_scale : 0.0f;
    }
    const int c0 = (ubase + u) * 16;
    uint8_t* prow = sP + (c0 >> 6) * 16384 + r * 128;
    const int u16 = (c0 & 63) >> 3;     // first of the two 16-byte units inside the 64-key chunk
    *reinterpret_cast<uint4*>(prow + (((u16) ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16(pr[0], pr[1]), pack_bf16(pr[2], pr[3]), pack_bf16(pr[4], pr[5]), pack_bf16(pr[6], pr[7]));
    *reinterpret_cast<uint4*>(prow + (((u16 + 1) ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16(pr[8], pr[9]), pack_bf16(pr[10], pr[11]), pack_bf16(pr[12], pr[13]), pack_bf16(pr[14], pr[15]));
  }
  // zero fill: the 16-column units of the band columns outside this row quadrant's window (split by parity)
#pragma unroll 1
  for (int uu = part; uu < NK / 16; uu += 2) {
    if (uu >= quad * 2 && uu < quad * 2 + 2 * UNITS) continue;   // warp-uniform
    const int c0 = uu * 16;
    uint8_t* prow = sP + (c0 >> 6) * 16384 + r * 128;
    const int u16 = (c0 & 63) >> 3;
    *reinterpret_cast<uint4*>(prow + ((u16 ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(prow + (((u16 + 1) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
  if (part == 1) {
    // global chunk: tile column NK holds the CLS key (absolute key 0), columns NK+1.. are zero
    float pg = exp2f(sg * LOG2E - m2);
    l += pg;
    if (p.drop_thresh != 0) pg *= attn_keep_cls(p.drop_seed, rowbase, p.drop_thresh, p.drop_scale);
    uint8_t* prow = sP + (NK >> 6) * 16384 + r * 128;
    const int u16 = (NK & 63) >> 3;
    *reinterpret_cast<uint4*>(prow + ((u16 ^ (r & 7)) << 4)) = make_uint4(pack_bf16(pg, 0.0f), 0, 0, 0);
    *reinterpret_cast<uint4*>(prow + (((u16 + 1) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }
  s_red[256 + part * 128 + r] = l;
  // ---- O = P V ----
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t ap = smem_u32(sP), av = smem_u32(sV);
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, HEAD_DIM, false, true);
#pragma unroll
    for (int ks = 0; ks < NT / 16; ++ks)
      umma_bf16(tmem, umma_smem_desc(ap + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                umma_smem_desc(av + ks * 2048, 8192, 1024), idesc2, ks > 0 ? 1u : 0u);
    umma_commit(bar_mma);
  }
  l = s_red[256 + r] + s_red[384 + r];
  const float inv_l = l > 0.0f ? 1.0f / l : 0.0f;
  const bool is_global_row = (i == 0) && (m_cls == 2);
  const bool do_store = (i < p.L) && !is_global_row;
  __nv_bfloat16* orow = p.ctx + (static_cast<size_t>(b) * p.L + (do_store ? i : 0)) * E + h * HEAD_DIM + part * 32;
  if (part == 0 && i < p.L)
    p.lse[(static_cast<size_t>(b) * p.H + h) * p.L + i] = (l > 0.0f) ? (m + logf(l)) : (p.lse_neg_inf ? -INFINITY : 0.0f);
  __syncwarp();
  mbar_wait(bar_mma, 1);
  tc_fence_after();
  {
    uint32_t v[32];
    tmem_ld32(lane_base + part * 32, v);   // warp-collective: executed by every lane
    tmem_ld_wait();
    if (do_store) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(v[j]) * inv_l, __uint_as_float(v[j + 1]) * inv_l);
        o.y = pack_bf16(__uint_as_float(v[j + 2]) * inv_l, __uint_as_float(v[j + 3]) * inv_l);
        o.z = pack_bf16(__uint_as_float(v[j + 4]) * inv_l, __uint_as_float(v[j + 5]) * inv_l);
        o.w = pack_bf16(__uint_as_float(v[j + 6]) * inv_l, __uint_as_float(v[j + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + j) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, C::TMEM_COLS);
  }
}

// ==============================================================================================
// attention_window 64 (W = 32): PERSISTENT kernel, one CTA (16 warps) per SM.
//
// Each CTA owns a contiguous run of the (b, h, tile) list (tile fastest, the same enumeration as the backward kernel,
// through the list of active query tiles when row activity is set).  What the one-shot kernel above pays per tile —
// CTA launch, barrier / TMEM set-up, ~3 700 cycles of exposed TMA latency, the S-MMA round trip — is taken off the
// critical path:
//   * operands (Q, K, V tiles; 68 KB) are double-buffered: tile t+1 is loaded while tile t is computed;
//   * S(t+1) = Q K^T is issued right behind the P V MMA of tile t, so it is ready when tile t+1 starts;
//   * four threads per query row (24 of the 96 window columns each), the key-valid bits built a tile ahead;
//   * P's out-of-window columns are zeroed ONCE (a thread writes the same window columns of its row every tile);
//   * the context tile is written with one 256-bit store per thread (a full 32-byte sector), the dropout keep bits
//     of a row (4 x u32: 24 bits per thread, the CLS column's bit in word 3) go to rf_attn_args.keepbits for the backward.
// ==============================================================================================
constexpr int AFP_THREADS = 512;
constexpr int AFP_W = 32, AFP_NK = 128 + 2 * AFP_W, AFP_NT = AFP_NK + 16;     // 192 band keys + the CLS chunk
constexpr uint32_t AFP_Q_BYTES = 128 * 128, AFP_KV_BYTES = AFP_NT * 128;
constexpr uint32_t AFP_STAGE = AFP_Q_BYTES + 2 * AFP_KV_BYTES;                 // 69 632 B: Q | K | V
constexpr uint32_t AFP_OFF_P = 2 * AFP_STAGE;
constexpr uint32_t AFP_OFF_BITS = AFP_OFF_P + 4 * 16384;
constexpr uint32_t AFP_OFF_RED = AFP_OFF_BITS + 64;                            // [2][4][128] floats: max / sum exchange
constexpr uint32_t AFP_OFF_BAR = AFP_OFF_RED + 2 * 4 * 128 * 4;
constexpr uint32_t AFP_SMEM = AFP_OFF_BAR + 64 + 1024;
static_assert(AFP_STAGE % 1024 == 0 && (AFP_Q_BYTES + AFP_KV_BYTES) % 1024 == 0, "swizzled tiles need 1024B alignment");
static_assert(AFP_SMEM <= 227 * 1024, "shared memory budget");

#ifdef RF_KTIMING
__device__ long long g_kt_fwd[4][16][16];     // [warp 0 / 5 / 10 / 15][tile][stamp]
#define KT(k) do { if (lane == 0 && (warp % 5) == 0 && blockIdx.x == 5 && it < 16) g_kt_fwd[warp / 5][it][k] = clock64(); } while (0)
#else
#define KT(k) do {} while (0)
#endif

struct AttnFwdPersistParams {
  AttnFwdParams a;
  uint32_t* keepbits;          // [B, H, L, 4] or null
  const int32_t* qtiles;       // rf_set_row_activity: compact list of active query tiles, or null = every tile
  const int32_t* n_qtiles;
};

__global__ void __launch_bounds__(AFP_THREADS)
band_attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
                             const AttnFwdPersistParams pp) {
  constexpr int W = AFP_W, NK = AFP_NK, NT = AFP_NT;
  const AttnFwdParams& p = pp.a;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sP = smem + AFP_OFF_P;
  uint32_t* kbits_all = reinterpret_cast<uint32_t*>(smem + AFP_OFF_BITS);      // [2][8]
  float* s_red = reinterpret_cast<float*>(smem + AFP_OFF_RED);
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + AFP_OFF_BAR);          // [2]: Q and K of a stage loaded
  uint64_t* bar_v = bar_qk + 2;                                                // [2]: V of a stage loaded
  uint64_t* bar_s = bar_qk + 4;                                                // S accumulator ready
  uint64_t* bar_o = bar_qk + 5;                                                // O accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int r = quad * 32 + lane;       // query row of the tile == TMEM lane
  const int E = p.H * HEAD_DIM;
  const int tiles_per_seq = (p.L + 127) / 128;
  const int n_q = pp.qtiles != nullptr ? *pp.n_qtiles : 0;
  const int total_tiles = pp.qtiles != nullptr ? n_q * p.H : p.B * p.H * tiles_per_seq;
  const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * total_tiles / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * total_tiles / gridDim.x);
  struct Pos { int tile, h, b, idx; };
  auto decode = [&](int t) -> Pos {
    Pos o;
    if (pp.qtiles != nullptr) {
      o.h = n_q > 0 ? t / n_q : 0;
      o.idx = t - o.h * n_q;
      const int q = pp.qtiles[o.idx];
      o.tile = q % tiles_per_seq;
      o.b = q / tiles_per_seq;
    } else {
      o.idx = 0;
      o.tile = t % tiles_per_seq;
      o.h = (t / tiles_per_seq) % p.H;
      o.b = t / (tiles_per_seq * p.H);
    }
    return o;
  };
  auto advance = [&](const Pos& c) -> Pos {
    Pos o = c;
    if (pp.qtiles != nullptr) {
      if (++o.idx == n_q) { o.idx = 0; ++o.h; }
      const int q = pp.qtiles[o.idx];
      o.tile = q % tiles_per_seq;
      o.b = q / tiles_per_seq;
    } else if (++o.tile == tiles_per_seq) {
      o.tile = 0;
      if (++o.h == p.H) { o.h = 0; ++o.b; }
    }
    return o;
  };
  // TMA loads of a tile's operands into stage `st`, in three groups issued by lane 0 of three different warps
  // (0: Q and the expect_tx arrival of the Q/K barrier, 1: K — its complete_tx may precede that arrival: the
  // transaction count goes negative, the phase cannot complete before the arrival —, 2: V on its own barrier)
  auto issue_loads = [&](const Pos& q, int st, int group) {
    uint8_t* sQ = smem + st * AFP_STAGE;
    uint8_t* sK = sQ + AFP_Q_BYTES;
    uint8_t* sV = sK + AFP_KV_BYTES;
    const int i0 = q.tile * 128, key0 = i0 - W + p.shift;
    if (group == 0) {
      mbar_arrive_expect_tx(&bar_qk[st], AFP_Q_BYTES + AFP_KV_BYTES);
#pragma unroll
      for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 8192, &tm64, &bar_qk[st], q.h * HEAD_DIM, i0 + c * 64, q.b);
    } else if (group == 1) {
#pragma unroll
      for (int c = 0; c < NK / 64; ++c) tma_load_3d(sK + c * 8192, &tm64, &bar_qk[st], E + q.h * HEAD_DIM, key0 + c * 64, q.b);
      tma_load_3d(sK + NK * 128, &tm16, &bar_qk[st], E + q.h * HEAD_DIM, 0, q.b);
    } else {
      mbar_arrive_expect_tx(&bar_v[st], AFP_KV_BYTES);
#pragma unroll
      for (int c = 0; c < NK / 64; ++c)
        tma_load_3d(sV + c * 8192, &tm64, &bar_v[st], 2 * E + q.h * HEAD_DIM, key0 + c * 64, q.b);
      tma_load_3d(sV + NK * 128, &tm16, &bar_v[st], 2 * E + q.h * HEAD_DIM, 0, q.b);
    }
  };
  // one mask byte per thread of warps 0..5 decides one key-valid bit of a tile (one unconditional load from a clamped
  // address: the byte is consumed later in the tile, nothing but the load sits at the issue point)
  auto kbyte_addr = [&](int b, int tile, bool& ok) -> const uint8_t* {
    const uint8_t* mrow = p.mask012 + static_cast<size_t>(b) * p.L;
    const int j = tile * 128 - W + p.shift + warp * 32 + lane;
    ok = warp < NK / 32 && j >= 0 && j < p.L;
    return mrow + (ok ? j : 0);
  };
  auto store_kbits = [&](uint32_t* kb, uint32_t kbyte, bool ok) {     // whole warps 0..5
    const uint32_t word = __ballot_sync(0xffffffffu, ok && kbyte == 1u);
    if (lane == 0) kb[warp] = word;
  };

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_qk[s], 1);
      mbar_init(&bar_v[s], 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  Pos cur = decode(t_begin < t_end ? t_begin : 0);
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0 && t_begin < t_end) {
    issue_loads(cur, 0, 0);
    issue_loads(cur, 0, 1);
    issue_loads(cur, 0, 2);
  }
  // P: zero the whole buffer once; every tile rewrites only the window columns of each row (and the CLS column)
  for (int o = tid; o < 4 * 16384 / 16; o += AFP_THREADS) *reinterpret_cast<uint4*>(sP + o * 16) = make_uint4(0, 0, 0, 0);
  if (t_begin < t_end && warp < NK / 32) {
    bool ok;
    const uint8_t* ka = kbyte_addr(cur.b, cur.tile, ok);
    store_kbits(kbits_all, *ka, ok);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t TM_S = 0, TM_O = 256;      // O is double-buffered: columns 256 + 64 * (tile parity)

  // Control warp 15 issues every MMA (warp-uniform descriptors, only the tcgen05 instructions on the elected lane)
  const bool ctrl = warp == 15;
  const bool elected = elect_one();
  auto issue_s = [&](int st) {      // S = Q K^T of the tile in stage `st`
    const uint32_t base = smem_u32(smem + st * AFP_STAGE);
    const uint64_t dq = umma_smem_desc(base, 16, 1024), dk = umma_smem_desc(base + AFP_Q_BYTES, 16, 1024);
    constexpr uint32_t idesc = umma_idesc_bf16(128, NT, false, false);
    if (elected) {
#pragma unroll
      for (int k = 0; k < HEAD_DIM / 16; ++k) umma_bf16(tmem + TM_S, dq + k * 2, dk + k * 2, idesc, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
  };
  if (ctrl && t_begin < t_end) {
    mbar_wait(&bar_qk[0], 0);
    tc_fence_after();
    issue_s(0);
  }
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const float LOG2E = 1.4426950408889634f;
  const int band_hi = 2 * W - p.hi_cut;   // last in-band column offset of a row
  const int c0 = quad * 32 + part * 24;   // first tile column of this thread's 24-column piece of the row's window
  const uint64_t dP_k = umma_smem_desc(smem_u32(sP), 16, 1024);

  // The context tile of tile t is stored during tile t+1, under that tile's P V MMA: what it needs is kept here
  float ep_inv_l = 0.f;
  __nv_bfloat16* ep_dst = nullptr;      // null = this thread stores nothing (row past L, or the global row)
  // (in the loop every thread has already waited for this accumulator before overwriting P; waiting again could miss
  //  the phase: by then the NEXT P V MMA may have completed and flipped the barrier's parity back)
  auto epilogue = [&](uint32_t itp, bool wait) {   // itp = local index of the tile whose accumulator is drained
    if (wait) mbar_wait(bar_o, itp & 1);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld16(lane_base + TM_O + (itp & 1) * 64 + part * 16, v);   // warp-collective: executed by every lane
    tmem_ld_wait();
    if (ep_dst != nullptr) {
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = pack_bf16(__uint_as_float(v[2 * j]) * ep_inv_l, __uint_as_float(v[2 * j + 1]) * ep_inv_l);
      st_global_v8(ep_dst, o);
    }
  };

  uint32_t it = 0;
#pragma unroll 1
  for (int t = t_begin; t < t_end; ++t, ++it) {
    const int st = it & 1;
    KT(0);
    const int tile = cur.tile, h = cur.h, b = cur.b;
    const bool has_next = t + 1 < t_end;
    Pos nxt = cur;
    if (has_next) nxt = advance(cur);
    // next tile: Q and K into the other stage (its previous tenant's S MMA retired a tile ago; V follows further down,
    // once the previous tile's P V MMA has retired), and the mask byte behind one of its key-valid bits
    if (has_next && quad == 3 && part < 2 && lane == 0) issue_loads(nxt, st ^ 1, part);
    __syncwarp();
    uint32_t kbyte_n = 0;
    bool kbyte_ok = false;
    if (has_next && warp < NK / 32)      // (volatile: must not be sunk down to its first use)
      asm volatile("ld.global.u8 %0, [%1];" : "=r"(kbyte_n) : "l"(kbyte_addr(nxt.b, nxt.tile, kbyte_ok)));
    const uint32_t* kbits = kbits_all + st * 8;
    const int i0 = tile * 128, i = i0 + r;
    const int key0 = i0 - W + p.shift;             // absolute key index of tile column 0
    const uint8_t* mrow = p.mask012 + static_cast<size_t>(b) * p.L;
    uint32_t m_row, m_cls;
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(m_row) : "l"(mrow + (i < p.L ? i : 0)));
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(m_cls) : "l"(mrow));
    const uint64_t rowid = (static_cast<uint64_t>(b) * p.H + h) * p.L + (i < p.L ? i : 0);
    const uint64_t rowbase = rowid * attn_drop_groups(p.L);

    KT(1);
    mbar_wait(bar_s, it & 1);
    tc_fence_after();
    KT(2);
    // ---- softmax: the four threads (quad, part 0..3) share query row i0 + r (TMEM lane r) ----
    uint32_t sv[24];
    float sg = -INFINITY;       // part 3: the CLS column's score
    {
      uint32_t a16[16], a8[8], g8[8];
      tmem_ld16(lane_base + TM_S + c0, a16);
      tmem_ld8(lane_base + TM_S + c0 + 16, a8);
      if (part == 3) tmem_ld8(lane_base + TM_S + NK, g8);   // warp-uniform
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) sv[j] = a16[j];
#pragma unroll
      for (int j = 0; j < 8; ++j) sv[16 + j] = a8[j];
      if (part == 3) sg = __uint_as_float(g8[0]);
    }
    const bool row_valid = (i < p.L) && (m_row != 0);
    const bool g_ok = p.use_cls && (m_cls == 2);
    uint32_t live;
    {
      const int wi = c0 >> 5, sh = c0 & 31;
      const uint64_t kw = (static_cast<uint64_t>(kbits[wi + 1]) << 32) | kbits[wi];     // word wi + 1 <= 6: masked off if unused
      const int lo = r - c0, hi = r + band_hi - c0;
      const uint32_t mlo = lo <= 0 ? 0xFFFFFFu : (lo >= 24 ? 0u : ((0xFFFFFFu << lo) & 0xFFFFFFu));
      const uint32_t mhi = hi >= 23 ? 0xFFFFFFu : (hi < 0 ? 0u : (0xFFFFFFu >> (23 - hi)));
      live = row_valid ? (static_cast<uint32_t>(kw >> sh) & mlo & mhi) : 0u;
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const float s = ((live >> j) & 1u) ? __uint_as_float(sv[j]) : -INFINITY;
      sv[j] = __float_as_uint(s);
      m = fmaxf(m, s);
    }
    if (!(part == 3 && g_ok && row_valid)) sg = -INFINITY;
    m = fmaxf(m, sg);
    s_red[part * 128 + r] = m;
    KT(3);
    // every thread has read its S columns; exchange the row maxima
    tc_fence_before();
    __syncthreads();
    KT(4);
    m = fmaxf(fmaxf(s_red[r], s_red[128 + r]), fmaxf(s_red[256 + r], s_red[384 + r]));
    if (m == -INFINITY) m = 0.0f;   // fully masked row: every p below is exp2(-inf) = 0
    const float m2 = m * LOG2E;

    float l = 0.0f;
    float pr[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      const float e = exp2f(__uint_as_float(sv[j]) * LOG2E - m2);
      l += e;
      pr[j] = e;
    }
    uint32_t keep = 0xFFFFFFu;
    if (p.drop_thresh != 0) {
      keep = live != 0 ? (attn_keep32(p.drop_seed, rowbase, key0 + c0, p.drop_thresh, live) & 0xFFFFFFu) : 0u;
#pragma unroll
      for (int j = 0; j < 24; ++j) pr[j] = ((keep >> j) & 1u) ? pr[j] * p.drop_scale : 0.0f;
    }
    float pg = 0.f;
    if (part == 3) {
      pg = exp2f(sg * LOG2E - m2);
      l += pg;
      if (p.drop_thresh != 0) {
        const float kg = attn_keep_cls(p.drop_seed, rowbase, p.drop_thresh, p.drop_scale);
        if (kg != 0.0f) keep |= 1u << 24;
        pg *= kg;
      }
    }
    // The previous tile's P V MMA must have retired before P is overwritten and before that tile's V buffer is
    // reloaded (it was issued a whole softmax ago: this wait is free); then V of the next tile
    if (it > 0) mbar_wait(bar_o, (it - 1) & 1);
    if (has_next && quad == 3 && part == 2 && lane == 0) issue_loads(nxt, st ^ 1, 2);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int c = c0 + u * 8;
      const uint32_t o = (c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(sP + o) = make_uint4(pack_bf16(pr[u * 8], pr[u * 8 + 1]), pack_bf16(pr[u * 8 + 2], pr[u * 8 + 3]),
                                                     pack_bf16(pr[u * 8 + 4], pr[u * 8 + 5]), pack_bf16(pr[u * 8 + 6], pr[u * 8 + 7]));
    }
    // global chunk: tile column NK holds the CLS key (absolute key 0), columns NK+1.. stay zero
    if (part == 3) *reinterpret_cast<uint4*>(sP + 3 * 16384 + r * 128 + ((r & 7) << 4)) = make_uint4(pack_bf16(pg, 0.0f), 0, 0, 0);
    if (pp.keepbits != nullptr && i < p.L) pp.keepbits[rowid * 4 + part] = keep;
    s_red[512 + part * 128 + r] = l;
    if (has_next && warp < NK / 32) {      // the next tile's key-valid bits (read after the barrier below)
      asm volatile("" : "+r"(kbyte_n));
      store_kbits(kbits_all + (st ^ 1) * 8, kbyte_n, kbyte_ok);
    }

    // ---- O = P V, then S of the next tile; meanwhile the previous tile's context rows are stored ----
    KT(5);
    fence_proxy_async_smem();
    __syncthreads();
    KT(6);
    if (ctrl) {
      mbar_wait(&bar_v[st], (it >> 1) & 1);
      tc_fence_after();
      const uint64_t dv = umma_smem_desc(smem_u32(smem + st * AFP_STAGE + AFP_Q_BYTES + AFP_KV_BYTES), 8192, 1024);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, HEAD_DIM, false, true);
      if (elected) {
#pragma unroll
        for (int ks = 0; ks < NT / 16; ++ks)
          umma_bf16(tmem + TM_O + st * 64, dP_k + ((ks >> 2) * 1024 + (ks & 3) * 2), dv + ks * 128, idesc2, ks > 0 ? 1u : 0u);
        umma_commit(bar_o);
      }
      __syncwarp();
      if (has_next) {
        mbar_wait(&bar_qk[st ^ 1], ((it + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(st ^ 1);
      }
    }
    KT(7);
    if (it > 0) epilogue(it - 1, false);
    KT(8);
    l = (s_red[512 + r] + s_red[640 + r]) + (s_red[768 + r] + s_red[896 + r]);
    ep_inv_l = l > 0.0f ? 1.0f / l : 0.0f;
    const bool is_global_row = (i == 0) && (m_cls == 2);
    ep_dst = ((i < p.L) && !is_global_row) ? p.ctx + (static_cast<size_t>(b) * p.L + i) * E + h * HEAD_DIM + part * 16 : nullptr;
    if (part == 0 && i < p.L)
      p.lse[rowid] = (l > 0.0f) ? (m + logf(l)) : (p.lse_neg_inf ? -INFINITY : 0.0f);
    cur = nxt;
    KT(9);
  }
  if (it > 0) epilogue(it - 1, true);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Running merge of window segments: acc / lse_acc hold the softmax-weighted output and log-sum-exp of the
// segments seen so far; (part, lse_part) is the next one.  grid over (token, head, 8-dim group).
__global__ void __launch_bounds__(256)
attn_merge_kernel(float* __restrict__ acc, float* __restrict__ lse_acc, const __nv_bfloat16* __restrict__ part,
                  const float* __restrict__ lse_part, const uint8_t* __restrict__ mask012, __nv_bfloat16* __restrict__ ctx,
                  float* __restrict__ lse_out, int B, int L, int H, int first, int last) {
  const long long idx = blockIdx.x * 256ll + threadIdx.x;
  const long long total = static_cast<long long>(B) * L * H * 8;
  if (idx >= total) return;
  const int g = static_cast<int>(idx & 7);
  const int h = static_cast<int>((idx >> 3) % H);
  const long long t = idx / (8ll * H);            // token index b * L + i
  const int b = static_cast<int>(t / L), i = static_cast<int>(t % L);
  if (i == 0 && mask012[static_cast<size_t>(b) * L] == 2) return;   // global row: written by rf_global_attn_fwd
  const int E = H * HEAD_DIM;
  const size_t o = static_cast<size_t>(t) * E + h * HEAD_DIM + g * 8;
  const size_t lo = (static_cast<size_t>(b) * H + h) * L + i;
  const float lp = lse_part[lo];
  const uint4 raw = *reinterpret_cast<const uint4*>(part + o);
  float v[8];
  {
    const float2 a0 = unpack_bf16(raw.x), a1 = unpack_bf16(raw.y), a2 = unpack_bf16(raw.z), a3 = unpack_bf16(raw.w);
    v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y; v[4] = a2.x; v[5] = a2.y; v[6] = a3.x; v[7] = a3.y;
  }
  float la = first ? -INFINITY : lse_acc[lo];
  const float mx = fmaxf(la, lp);
  float wa = 0.f, wp = 0.f, ln = -INFINITY;
  if (mx > -INFINITY) {
    const float ea = __expf(la - mx), ep = __expf(lp - mx);
    ln = mx + __logf(ea + ep);
    wa = ea / (ea + ep); wp = ep / (ea + ep);
  }
  float r[8];
  if (first) {
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = wp * v[e];
  } else {
    const float4 c0 = *reinterpret_cast<const float4*>(acc + o), c1 = *reinterpret_cast<const float4*>(acc + o + 4);
    const float a[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = wa * a[e] + wp * v[e];
  }
  if (last) {
    uint4 out;
    out.x = pack_bf16(r[0], r[1]); out.y = pack_bf16(r[2], r[3]); out.z = pack_bf16(r[4], r[5]); out.w = pack_bf16(r[6], r[7]);
    *reinterpret_cast<uint4*>(ctx + o) = out;
    if (g == 0) lse_out[lo] = (ln > -INFINITY) ? ln : 0.0f;
  } else {
    *reinterpret_cast<float4*>(acc + o) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(acc + o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    if (g == 0) lse_acc[lo] = ln;
  }
}

struct AttnSegment { int shift, hi_cut, use_cls; };

// Native half-width of the forward kernel used for a band of half-width w: the widest of {32, 64, 128} that
// does not exceed w.
static int fwd_native_w(int w) { return w >= 128 ? 128 : (w >= 64 ? 64 : 32); }

// Segments of a band of half-width w for a kernel of native half-width wk (2*wk+1 key offsets each): offsets
// [-w + (2wk+1)k, -w + (2wk+1)k + 2wk] clipped to [-w, w]; the kernel's own band is centred, so segment k is run
// with keys shifted by its centre.
static int attn_segments(int w, int wk, AttnSegment* seg) {
  const int span = 2 * wk + 1;
  const int n = (2 * w + 1 + span - 1) / span;
  for (int k = 0; k < n; ++k) {
    const int lo = -w + span * k, hi = lo + 2 * wk;
    seg[k].shift = lo + wk;
    seg[k].hi_cut = hi > w ? hi - w : 0;
    seg[k].use_cls = k == 0;
  }
  return n;
}

template <int W>
static int launch_attn_fwd(const rf_attn_args* a, void* ctx, float* lse, const AttnSegment& sg, int lse_neg_inf,
                           cudaStream_t stream) {
  using C = AttnFwdCfg<W>;
  auto kern = band_attn_fwd_kernel<W>;
  static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
  if (first_use_on_device(&attr_seen)) {
    RF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
  }
  const int E = a->H * HEAD_DIM;
  const CUtensorMap* tm64 = get_tmap_3d(a->qkv, a->B, a->L, 3 * E, 3 * E, static_cast<uint64_t>(a->L) * 3 * E, 64);
  const CUtensorMap* tm16 = get_tmap_3d(a->qkv, a->B, a->L, 3 * E, 3 * E, static_cast<uint64_t>(a->L) * 3 * E, 16);
  if (!tm64 || !tm16) return RF_ERR_CUDA;
  AttnFwdParams p;
  p.mask012 = a->mask012;
  p.ctx = reinterpret_cast<__nv_bfloat16*>(ctx);
  p.lse = lse;
  p.B = a->B; p.L = a->L; p.H = a->H;
  p.shift = sg.shift; p.hi_cut = sg.hi_cut; p.use_cls = sg.use_cls; p.lse_neg_inf = lse_neg_inf;
  p.drop_thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  p.drop_scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  p.drop_seed = a->drop_seed;
  {
    const RowActivity& ra = row_activity();
    p.row_active = (ra.flags != nullptr && ra.rows == static_cast<long long>(a->B) * a->L && a->L % 256 == 0) ? ra.flags
                                                                                                              : nullptr;
  }
  const int tiles = (a->L + 127) / 128;
  kern<<<a->B * a->H * tiles, ATT_THREADS, C::TOTAL, stream>>>(*tm64, *tm16, p);
  return check_launch("rf_band_attn_fwd");
}

// attention_window 64: the persistent kernel
static int launch_attn_fwd_persist(const rf_attn_args* a, void* ctx, float* lse, const AttnSegment& sg, int lse_neg_inf,
                                   cudaStream_t stream) {
  static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
  if (first_use_on_device(&attr_seen)) {
    RF_CUDA(cudaFuncSetAttribute(band_attn_fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AFP_SMEM));
  }
  RF_REQUIRE((reinterpret_cast<uintptr_t>(ctx) & 31) == 0, "rf_band_attn_fwd: ctx must be 32-byte aligned (256-bit stores)");
  const int E = a->H * HEAD_DIM;
  const CUtensorMap* tm64 = get_tmap_3d(a->qkv, a->B, a->L, 3 * E, 3 * E, static_cast<uint64_t>(a->L) * 3 * E, 64);
  const CUtensorMap* tm16 = get_tmap_3d(a->qkv, a->B, a->L, 3 * E, 3 * E, static_cast<uint64_t>(a->L) * 3 * E, 16);
  if (!tm64 || !tm16) return RF_ERR_CUDA;
  AttnFwdPersistParams pp;
  AttnFwdParams& p = pp.a;
  p.mask012 = a->mask012;
  p.ctx = reinterpret_cast<__nv_bfloat16*>(ctx);
  p.lse = lse;
  p.B = a->B; p.L = a->L; p.H = a->H;
  p.shift = sg.shift; p.hi_cut = sg.hi_cut; p.use_cls = sg.use_cls; p.lse_neg_inf = lse_neg_inf;
  p.drop_thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  p.drop_scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  p.drop_seed = a->drop_seed;
  p.row_active = nullptr;
  pp.keepbits = (a->w == 32 && a->drop_p > 0.f) ? reinterpret_cast<uint32_t*>(a->keepbits) : nullptr;
  const RowActivity& ra = row_activity();
  const bool on = ra.flags != nullptr && ra.rows == static_cast<long long>(a->B) * a->L && a->L % 256 == 0;
  pp.qtiles = on ? ra.qtiles : nullptr;
  pp.n_qtiles = on ? ra.n_qtiles : nullptr;
  const int total = a->B * a->H * ((a->L + 127) / 128);
  band_attn_fwd_persist_kernel<<<total < sm_count() ? total : sm_count(), AFP_THREADS, AFP_SMEM, stream>>>(*tm64, *tm16, pp);
  return check_launch("rf_band_attn_fwd");
}

static int launch_attn_fwd_w(int wk, const rf_attn_args* a, void* ctx, float* lse, const AttnSegment& sg, int lse_neg_inf,
                             cudaStream_t stream) {
  if (wk == 32) return launch_attn_fwd_persist(a, ctx, lse, sg, lse_neg_inf, stream);
  if (wk == 128) return launch_attn_fwd<128>(a, ctx, lse, sg, lse_neg_inf, stream);
  if (wk == 64) return launch_attn_fwd<64>(a, ctx, lse, sg, lse_neg_inf, stream);
  return set_error(RF_ERR_INVALID, "rf_band_attn_fwd: no kernel for native half-width %d", wk);
}

}  // namespace rf

#ifdef RF_KTIMING
extern "C" int rf_debug_ktiming_fwd(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, rf::g_kt_fwd, sizeof(rf::g_kt_fwd)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" long long rf_band_attn_ws_bytes(int B, int L, int H, int w) {
  if (w <= 32) return 0;
  const long long T = static_cast<long long>(B) * L, E = static_cast<long long>(H) * rf::HEAD_DIM;
  // forward (segmented windows only): fp32 accumulator [T,E] + bf16 segment output [T,E] + 2 x lse [B,H,L];
  // backward: fp32 dQ scratch [T,E] (aliases the forward accumulator)
  return T * E * 4 + T * E * 2 + 2ll * B * H * L * 4 + 256;
}

extern "C" int rf_band_attn_fwd(const rf_attn_args* a, void* ctx, float* lse, rf_stream_t stream_) {
  using namespace rf;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && ctx && lse, "rf_band_attn_fwd: null argument");
  RF_REQUIRE(a->D == HEAD_DIM, "rf_band_attn_fwd: head_dim %d unsupported (64 only)", a->D);
  RF_REQUIRE(a->B > 0 && a->L >= 16 && a->H > 0, "rf_band_attn_fwd: bad shape B=%d L=%d H=%d", a->B, a->L, a->H);
  RF_REQUIRE(a->w >= 32 && a->w % 32 == 0 && a->w <= 256,
             "rf_band_attn_fwd: one-sided window %d unsupported (multiples of 32 up to 256)", a->w);
  RF_REQUIRE(a->L <= ATTN_MAX_L, "rf_band_attn_fwd: sequence length %d exceeds %d", a->L, ATTN_MAX_L);
  AttnSegment seg[16];
  const int wk = fwd_native_w(a->w);
  const int nseg = attn_segments(a->w, wk, seg);
  if (nseg == 1) return launch_attn_fwd_w(wk, a, ctx, lse, seg[0], 0, stream);
  RF_REQUIRE(a->ws != nullptr, "rf_band_attn_fwd: this window needs a workspace (rf_band_attn_ws_bytes)");
  const size_t T = static_cast<size_t>(a->B) * a->L, E = static_cast<size_t>(a->H) * HEAD_DIM;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->ws);
  float* acc = reinterpret_cast<float*>(ws);
  __nv_bfloat16* part = reinterpret_cast<__nv_bfloat16*>(ws + T * E * 4);
  float* lse_part = reinterpret_cast<float*>(ws + T * E * 6);
  float* lse_acc = lse_part + static_cast<size_t>(a->B) * a->H * a->L;
  const long long total = static_cast<long long>(T) * a->H * 8;
  for (int k = 0; k < nseg; ++k) {
    int rc = launch_attn_fwd_w(wk, a, part, lse_part, seg[k], 1, stream);
    if (rc) return rc;
    attn_merge_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        acc, lse_acc, part, lse_part, a->mask012, reinterpret_cast<__nv_bfloat16*>(ctx), lse, a->B, a->L, a->H, k == 0,
        k == nseg - 1);
    rc = check_launch("rf_band_attn_fwd/merge");
    if (rc) return rc;
  }
  return RF_OK;
}
