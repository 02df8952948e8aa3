// Banded attention backward, sm_100a.  The kernel covers 65 key offsets per query (one-sided window 32 =
// attention_window 64); wider windows run it once per window segment (shifted keys, see attention_fwd.cu)
// with the FINAL log-sum-exp / context of the merged forward, accumulating dQ in an fp32 scratch.
//
// One CTA = one (batch, head, 128-query tile), the same tiling as the forward kernel:
//   S  = Q K^T,  dP = dO V^T                      tcgen05.mma -> TMEM (2 x 208 columns)
//   delta = rowsum(dO o O) (= sum_j P'_j dP_j, taken from the saved context instead of a first pass
//   over the probabilities),  P = exp(S - lse),  dS = P (dP - delta)    fp32, thread = query row
//   P, dS (bf16) -> shared memory (128B-swizzled, rows = queries)
//   dQ = dS K          A = dS (K-major),   B = K  (MN-major view of the K tile)
//   dV = P^T dO        A = P  (MN-major view of the same P buffer), B = dO (MN-major)
//   dK = dS^T Q        A = dS (MN-major view),                       B = Q  (MN-major)
// so no operand is ever transposed in memory.  dQ (scaled back by 1/sqrt(D)) is written as bf16
// into dqkv.  dK/dV tiles overlap between neighbouring query tiles: at attention_window 64 a key receives at most
// two partial sums, which go straight into the bf16 gradient with 16-byte red.add.noftz.v4.bf16x2 (only the CLS
// key, which every tile feeds, is accumulated in fp32 and folded in by a tiny kernel); wide windows (several
// segments per key) accumulate dK/dV and dQ in fp32 scratch that a fold kernel converts.  The global query row
// receives no band gradient (its band output is overwritten by the global row, HF:615-626).
// Key validity is a bit mask per tile (band position by shifts: no per-element shared-memory flag loads), and the
// dropout masks are regenerated from ABSOLUTE (row, key) coordinates (rf_ptx.cuh), independent of the tiling.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

RF_DEFINE_NONCE_LOADER(attn_bwd)

namespace rf {

constexpr int AB_THREADS = 512;
constexpr int AB_W = 32;
constexpr int AB_NK = 128 + 2 * AB_W;   // 192
constexpr int AB_NT = AB_NK + 16;       // 208
constexpr int AB_D = 64;
constexpr uint32_t AB_Q_BYTES = 128 * 128, AB_KV_BYTES = AB_NT * 128, AB_P_BYTES = 4 * 16384;
constexpr uint32_t AB_OFF_Q = 0;
constexpr uint32_t AB_OFF_DO = AB_OFF_Q + AB_Q_BYTES;
constexpr uint32_t AB_OFF_K = AB_OFF_DO + AB_Q_BYTES;
constexpr uint32_t AB_OFF_V = AB_OFF_K + AB_KV_BYTES;          // 26624 = 26 KiB: stays 1024-aligned
constexpr uint32_t AB_OFF_P = AB_OFF_V + AB_KV_BYTES;
constexpr uint32_t AB_OFF_DS = AB_OFF_P + AB_P_BYTES;
constexpr uint32_t AB_OFF_FLAG = AB_OFF_DS + AB_P_BYTES;
constexpr uint32_t AB_OFF_BAR = AB_OFF_FLAG + 64;     // key-valid bit words (7 used)
constexpr uint32_t AB_OFF_DELTA = AB_OFF_BAR + 64;
constexpr uint32_t AB_SMEM = AB_OFF_DELTA + 4 * 128 * 4 + 1024;
static_assert(AB_OFF_V % 1024 == 0 && AB_OFF_P % 1024 == 0, "swizzled tiles need 1024B alignment");
static_assert(AB_SMEM <= 227 * 1024, "shared memory budget");

struct AttnBwdParams {
  const uint8_t* mask012;
  const __nv_bfloat16* ctx;   // forward output O (bf16 [B*L, E])
  const float* lse;
  __nv_bfloat16* dqkv;
  float* dkv;   // fp32 scratch [B*L, 2E] (wide windows: many partial sums per key), or null:
  float* dkv_cls;   // attention_window 64: dK/dV go straight into dqkv as bf16x2 red.adds (a key receives at most two
                    // partial sums); only the CLS key, which every tile feeds, is accumulated in fp32 here [B,H,2,64]
  float* dq32;  // fp32 dQ scratch [B*L, E] (wide windows: dQ is accumulated over the window segments) or null
  int B, L, H;
  int shift, hi_cut, use_cls;   // window segment (see attention_fwd.cu)
  float drop_scale;
  uint32_t drop_thresh;
  uint64_t drop_seed;
  // row activity (rf_set_row_activity): the persistent CTAs walk the compact list of active query tiles x heads
  const int32_t* qtiles;      // [n] entries b * tiles_per_seq + tile, or null = every tile
  const int32_t* n_qtiles;    // device scalar n
};

__global__ void __launch_bounds__(AB_THREADS)
band_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV64, const __grid_constant__ CUtensorMap tmQKV16,
                     const __grid_constant__ CUtensorMap tmDO, const AttnBwdParams p) {
  constexpr int W = AB_W, NK = AB_NK, NT = AB_NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem + AB_OFF_Q;
  uint8_t* sDO = smem + AB_OFF_DO;
  uint8_t* sK = smem + AB_OFF_K;
  uint8_t* sV = smem + AB_OFF_V;
  uint8_t* sP = smem + AB_OFF_P;
  uint8_t* sDS = smem + AB_OFF_DS;
  uint32_t* kbits = reinterpret_cast<uint32_t*>(smem + AB_OFF_FLAG);
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + AB_OFF_BAR);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);
  uint64_t* bar_kv = bar_load + 3;     // dK / dV accumulators ready (dQ is published earlier on bar_mma)
  float* s_delta = reinterpret_cast<float*>(smem + AB_OFF_DELTA);   // [4][128] partial row sums

  // 16 warps: TMEM lane quadrant = warp % 4 (rows 32*quad..), `part` = warp / 4 splits every row's work
  // between four threads (the kernel runs one CTA per SM: its warps are all the latency hiding there is):
  //   softmax backward: parts 0..2 take one 32-column window chunk each, part 3 the CLS column + the zero fill
  //   delta / dQ: 16 of the 64 head dims each;  dK / dV: part = (dK | dV, low | high 32 dims)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int tiles_per_seq = (p.L + 127) / 128;
  const int total_tiles = p.qtiles != nullptr ? *p.n_qtiles * p.H : p.B * p.H * tiles_per_seq;
  // work item t -> (sequence b, head h, query tile): dense enumeration, or through the list of active query tiles
  auto decode = [&](int t, int& tile, int& h, int& b) {
    if (p.qtiles != nullptr) {
      const int q = p.qtiles[t / p.H];
      h = t % p.H;
      tile = q % tiles_per_seq;
      b = q / tiles_per_seq;
    } else {
      tile = t % tiles_per_seq;
      h = (t / tiles_per_seq) % p.H;
      b = t / (tiles_per_seq * p.H);
    }
  };
  const int E = p.H * AB_D;
  constexpr uint32_t TM_S = 0, TM_DP = 256;                       // phase 1
  constexpr uint32_t TM_DQ = 0, TM_DV = 64, TM_DK = 192;          // phase 2 (dV: 2 x 64, dK: 2 x 64)

  // PERSISTENT: one CTA per SM walks the (b, h, tile) list.  The operand loads of the NEXT tile are issued as soon as
  // the current tile's last MMAs have retired (every shared-memory operand is dead then), so they land while the dK / dV
  // epilogue drains TMEM; barriers and the 512 TMEM columns are set up once per CTA.
  auto issue_loads = [&](int t) {     // thread 0 only
    int tile, h, b;
    decode(t, tile, h, b);
    const int i0 = tile * 128, key0 = i0 - W + p.shift;
    mbar_arrive_expect_tx(bar_load, 2 * AB_Q_BYTES + 2 * AB_KV_BYTES);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tma_load_3d(sQ + c * 8192, &tmQKV64, bar_load, h * AB_D, i0 + c * 64, b);
      tma_load_3d(sDO + c * 8192, &tmDO, bar_load, h * AB_D, i0 + c * 64, b);
    }
#pragma unroll
    for (int c = 0; c < NK / 64; ++c) {
      tma_load_3d(sK + c * 8192, &tmQKV64, bar_load, E + h * AB_D, key0 + c * 64, b);
      tma_load_3d(sV + c * 8192, &tmQKV64, bar_load, 2 * E + h * AB_D, key0 + c * 64, b);
    }
    tma_load_3d(sK + NK * 128, &tmQKV16, bar_load, E + h * AB_D, 0, b);
    tma_load_3d(sV + NK * 128, &tmQKV16, bar_load, 2 * E + h * AB_D, 0, b);
  };
  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_kv, 1);
    fence_mbar_init();
    if (static_cast<int>(blockIdx.x) < total_tiles) issue_loads(blockIdx.x);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  uint32_t it = 0;      // tiles processed by this CTA: parity of the once-per-tile barriers
#pragma unroll 1
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
  int tile, h, b;
  decode(t, tile, h, b);
  const int i0 = tile * 128;
  const uint8_t* mrow = p.mask012 + static_cast<size_t>(b) * p.L;
  const int key0 = i0 - W + p.shift;             // absolute key index of tile column 0
  // Per-row global loads issued right away (they are first used after the S / dP MMAs): this thread's quarter
  // (16 of 64 dims) of the saved context row, the row's log-sum-exp and its mask bytes.
  const int r = quad * 32 + lane;       // query row of the tile == TMEM lane
  const int i = i0 + r;
  const bool in_seq = i < p.L;
  uint4 o_raw[2];
  {
    const uint4* op = reinterpret_cast<const uint4*>(p.ctx + (static_cast<size_t>(b) * p.L + (in_seq ? i : 0)) * E +
                                                     h * AB_D + part * 16);
#pragma unroll
    for (int u = 0; u < 2; ++u) o_raw[u] = in_seq ? op[u] : make_uint4(0, 0, 0, 0);
  }
  const float lse = in_seq ? p.lse[(static_cast<size_t>(b) * p.H + h) * p.L + i] : 0.f;
  const uint8_t m_row = mrow[in_seq ? i : 0], m_cls = mrow[0];
  // key-valid bits of the NK band columns (bit c: key is in range, not padding, not global); word 7 bit 0 = CLS column
  if (warp < NK / 32) {
    const int j = key0 + warp * 32 + lane;
    const bool inr = j >= 0 && j < p.L;
    const uint32_t word = __ballot_sync(0xffffffffu, inr && (mrow[inr ? j : 0] == 1));
    if (lane == 0) kbits[warp] = word;
  } else if (warp == NK / 32 && lane == 0) {
    kbits[7] = (p.use_cls && mrow[0] == 2) ? 1u : 0u;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_wait(bar_load, it & 1);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, NT, false, false);
    const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), ado = smem_u32(sDO), av = smem_u32(sV);
#pragma unroll
    for (int k = 0; k < AB_D / 16; ++k)
      umma_bf16(tmem + TM_S, umma_smem_desc(aq + k * 32, 16, 1024), umma_smem_desc(ak + k * 32, 16, 1024), idesc,
                k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < AB_D / 16; ++k)
      umma_bf16(tmem + TM_DP, umma_smem_desc(ado + k * 32, 16, 1024), umma_smem_desc(av + k * 32, 16, 1024), idesc,
                k > 0 ? 1u : 0u);
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  const bool is_global_row = (i == 0) && (m_cls == 2);
  const bool row_valid = in_seq && (m_row != 0) && !is_global_row;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const float LOG2E = 1.4426950408889634f;
  const float lse2 = lse * LOG2E;
  const uint64_t rowid = (static_cast<uint64_t>(b) * p.H + h) * p.L + (in_seq ? i : 0);
  const uint64_t rowbase = rowid * attn_drop_groups(p.L);
  const bool g_ok = kbits[7] != 0;
  // window chunks of this row block: quad + part for parts 0..2; part 3 has the CLS column and the zero chunks
  const int cc_lo = quad + part;
  const int cc_hi = part < 3 ? quad + part + 1 : cc_lo;

  // live columns of a 32-column chunk for this row: key-valid bits AND band position [r, r + 2W - hi_cut]
  const int band_hi = 2 * W - p.hi_cut;
  auto chunk_live = [&](int cc) -> uint32_t {
    const int lo = r - cc * 32, hi = r + band_hi - cc * 32;
    const uint32_t mlo = lo <= 0 ? 0xFFFFFFFFu : (lo >= 32 ? 0u : (0xFFFFFFFFu << lo));
    const uint32_t mhi = hi >= 31 ? 0xFFFFFFFFu : (hi < 0 ? 0u : (0xFFFFFFFFu >> (31 - hi)));
    return row_valid ? (kbits[cc] & mlo & mhi) : 0u;
  };

  // ---- delta_i = sum_j P'_ij dP_ij = dO_i . O_i  (O = sum_j P'_ij V_j is the saved forward output):
  //      each part dots its 16 dims (dO from the swizzled shared-memory tile, O from registers) ----
  float delta = 0.f;
  {
    const uint8_t* drow = sDO + (r >> 6) * 8192 + (r & 63) * 128;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint4 d = *reinterpret_cast<const uint4*>(drow + (((part * 2 + u) ^ (r & 7)) << 4));
      const uint4 o = o_raw[u];
      const float2 d0 = unpack_bf16(d.x), d1 = unpack_bf16(d.y), d2 = unpack_bf16(d.z), d3 = unpack_bf16(d.w);
      const float2 o0 = unpack_bf16(o.x), o1 = unpack_bf16(o.y), o2 = unpack_bf16(o.z), o3 = unpack_bf16(o.w);
      delta += d0.x * o0.x + d0.y * o0.y + d1.x * o1.x + d1.y * o1.y + d2.x * o2.x + d2.y * o2.y + d3.x * o3.x +
               d3.y * o3.y;
    }
  }
  float pg = 0.f, pg_d = 0.f, keep_g = 1.f, dpg = 0.f;
  if (part == 3) {   // warp-uniform
    uint32_t gs[16], gd[16];
    tmem_ld16(lane_base + TM_S + NK, gs);
    tmem_ld16(lane_base + TM_DP + NK, gd);
    tmem_ld_wait();
    pg = (row_valid && g_ok) ? exp2f(__uint_as_float(gs[0]) * LOG2E - lse2) : 0.f;   // undropped
    pg_d = pg;
    if (p.drop_thresh != 0) {
      keep_g = attn_keep_cls(p.drop_seed, rowbase, p.drop_thresh, p.drop_scale);
      pg_d = pg * keep_g;
    }
    dpg = __uint_as_float(gd[0]);
  }
  s_delta[part * 128 + r] = delta;
  __syncthreads();
  // masked rows may hold non-finite garbage in O / dO
  delta = row_valid ? (s_delta[r] + s_delta[128 + r]) + (s_delta[256 + r] + s_delta[384 + r]) : 0.f;

  // ---- pass B: P' and dS -> shared memory (one window chunk per part 0..2; part 3 fills the all-zero chunks) ----
#pragma unroll 1
  for (int cc = 0; cc < NK / 32; ++cc) {
    const bool in_win = cc >= quad && cc < quad + 3;
    const bool mine = in_win ? (cc >= cc_lo && cc < cc_hi) : (part == 3);
    if (!mine) continue;   // warp-uniform
    uint4 po[4], so[4];
    if (in_win) {
      uint32_t sv[32], dv[32];
      tmem_ld32(lane_base + TM_S + cc * 32, sv);
      tmem_ld32(lane_base + TM_DP + cc * 32, dv);
      tmem_ld_wait();
      const uint32_t live = chunk_live(cc);
      const uint32_t keepm = (p.drop_thresh != 0 && live != 0)
                                 ? attn_keep32(p.drop_seed, rowbase, key0 + cc * 32, p.drop_thresh, live) : 0xFFFFFFFFu;
      float pr[32], ds[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pu = ((live >> j) & 1u) ? exp2f(__uint_as_float(sv[j]) * LOG2E - lse2) : 0.f;
        const float kp = ((keepm >> j) & 1u) ? p.drop_scale : 0.f;
        pr[j] = pu * kp;                                           // P' feeds dV
        ds[j] = pu * (kp * __uint_as_float(dv[j]) - delta);        // softmax backward
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        po[u] = make_uint4(pack_bf16(pr[u * 8], pr[u * 8 + 1]), pack_bf16(pr[u * 8 + 2], pr[u * 8 + 3]),
                           pack_bf16(pr[u * 8 + 4], pr[u * 8 + 5]), pack_bf16(pr[u * 8 + 6], pr[u * 8 + 7]));
        so[u] = make_uint4(pack_bf16(ds[u * 8], ds[u * 8 + 1]), pack_bf16(ds[u * 8 + 2], ds[u * 8 + 3]),
                           pack_bf16(ds[u * 8 + 4], ds[u * 8 + 5]), pack_bf16(ds[u * 8 + 6], ds[u * 8 + 7]));
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) po[u] = so[u] = make_uint4(0, 0, 0, 0);
    }
    const uint32_t roff = (cc >> 1) * 16384 + r * 128;
    const int ubase = (cc & 1) * 4;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t o = roff + (((ubase + u) ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(sP + o) = po[u];
      *reinterpret_cast<uint4*>(sDS + o) = so[u];
    }
  }
  {
    // global chunk = P/dS chunk 3 (columns 192..255): column 192 holds the CLS key, the rest is zero;
    // part 3 (which owns pg) writes units 0..3, part 0 units 4..7
    const float dsg = pg * (keep_g * dpg - delta);
    const uint32_t roff = 3 * 16384 + r * 128;
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
      if (part == 1 || part == 2) continue;   // warp-uniform
      const int u = part == 3 ? uu : uu + 4;
      const uint32_t o = roff + ((u ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(sP + o) = (u == 0) ? make_uint4(pack_bf16(pg_d, 0.f), 0, 0, 0) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sDS + o) = (u == 0) ? make_uint4(pack_bf16(dsg, 0.f), 0, 0, 0) : make_uint4(0, 0, 0, 0);
    }
  }

  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), ado = smem_u32(sDO), ap = smem_u32(sP), ads = smem_u32(sDS);
    // dQ[128 x 64] = dS[128 x 208] K[208 x 64]
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, AB_D, false, true);
#pragma unroll
    for (int ks = 0; ks < NT / 16; ++ks)
      umma_bf16(tmem + TM_DQ, umma_smem_desc(ads + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                umma_smem_desc(ak + ks * 2048, 8192, 1024), idesc_q, ks > 0 ? 1u : 0u);
    umma_commit(bar_mma);   // dQ is stored while the dK / dV MMAs below still run
    // dV[keys x 64] = P^T dO,  dK[keys x 64] = dS^T Q   (two 128-key halves each)
    constexpr uint32_t idesc_kv = umma_idesc_bf16(128, AB_D, true, true);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
      for (int ks = 0; ks < 128 / 16; ++ks) {
        umma_bf16(tmem + TM_DV + hh * 64, umma_smem_desc(ap + hh * 32768 + ks * 2048, 16384, 1024),
                  umma_smem_desc(ado + ks * 2048, 8192, 1024), idesc_kv, ks > 0 ? 1u : 0u);
        umma_bf16(tmem + TM_DK + hh * 64, umma_smem_desc(ads + hh * 32768 + ks * 2048, 16384, 1024),
                  umma_smem_desc(aq + ks * 2048, 8192, 1024), idesc_kv, ks > 0 ? 1u : 0u);
      }
    }
    umma_commit(bar_kv);
  }
  __syncwarp();
  mbar_wait(bar_mma, 1);
  tc_fence_after();

  // ---- dQ (x 1/sqrt(D): gradient w.r.t. the unscaled projection); part p owns columns 16p..16p+15 ----
  {
    uint32_t v[16];
    tmem_ld16(lane_base + TM_DQ + part * 16, v);
    tmem_ld_wait();
    if (in_seq && p.dq32 != nullptr) {
      float* orow = p.dq32 + (static_cast<size_t>(b) * p.L + i) * E + h * AB_D + part * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + j), "f"(__uint_as_float(v[j]) * 0.125f),
                     "f"(__uint_as_float(v[j + 1]) * 0.125f), "f"(__uint_as_float(v[j + 2]) * 0.125f),
                     "f"(__uint_as_float(v[j + 3]) * 0.125f)
                     : "memory");
    } else if (in_seq) {
      __nv_bfloat16* orow = p.dqkv + (static_cast<size_t>(b) * p.L + i) * 3 * E + h * AB_D + part * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(v[j]) * 0.125f, __uint_as_float(v[j + 1]) * 0.125f);
        o.y = pack_bf16(__uint_as_float(v[j + 2]) * 0.125f, __uint_as_float(v[j + 3]) * 0.125f);
        o.z = pack_bf16(__uint_as_float(v[j + 4]) * 0.125f, __uint_as_float(v[j + 5]) * 0.125f);
        o.w = pack_bf16(__uint_as_float(v[j + 6]) * 0.125f, __uint_as_float(v[j + 7]) * 0.125f);
        *reinterpret_cast<uint4*>(orow + j) = o;
      }
    }
  }
  mbar_wait(bar_kv, it & 1);
  tc_fence_after();
  // every MMA of this tile has retired: Q / dO / K / V are dead -> start the next tile's loads under the epilogue below
  if (t + static_cast<int>(gridDim.x) < total_tiles) {
    const int tn = t + gridDim.x;
    if (tid == 0) issue_loads(tn);
    // ... and pull the next tile's per-row data (saved context quarter, log-sum-exp) into L2: those plain loads sit on
    // the next iteration's critical path (their DRAM latency was the top stall of the one-shot kernel)
    int tilen, hn, bn;
    decode(tn, tilen, hn, bn);
    const int in_ = tilen * 128 + r;
    if (in_ < p.L) {
      const void* pc = p.ctx + (static_cast<size_t>(bn) * p.L + in_) * E + hn * AB_D + part * 16;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pc));
      if (part == 0) {
        const void* pl = p.lse + (static_cast<size_t>(bn) * p.H + hn) * p.L + in_;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pl));
      }
    }
  }
  // every shared-memory operand is dead now: the P region becomes 16 per-warp 4 KB transpose slabs
  uint8_t* slab = sP + warp * 4096;
  // ---- dK / dV: TMEM lane = key column c = hh*128 + r of the tile.  Each 32-key x 32-dim chunk is
  //      transposed through the warp's slab so that one red.add.v4 instruction covers 4 key rows x
  //      128 contiguous bytes (4 LSU wavefronts) instead of 32 rows x 16 bytes. ----
  const int which = part >> 1, dhalf = part & 1;   // this warp: dK (0) or dV (1), dims 32*dhalf .. +31
#pragma unroll 1
  for (int hh = 0; hh < 2; ++hh) {    // the two 128-key halves of the tile
    const uint32_t tcol = (which == 0 ? TM_DK : TM_DV) + hh * 64 + dhalf * 32;
    uint32_t v[32];
    tmem_ld32(lane_base + tcol, v);
    tmem_ld_wait();
    {
      uint8_t* srow = slab + lane * 128;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<uint4*>(srow + ((u ^ (lane & 7)) << 4)) = make_uint4(v[u * 4], v[u * 4 + 1], v[u * 4 + 2], v[u * 4 + 3]);
    }
    __syncwarp();
    if (p.dkv != nullptr) {
#pragma unroll
      for (int s2 = 0; s2 < 8; ++s2) {
        const int rl = s2 * 4 + (lane >> 3);          // key row within this warp's 32
        const int u = lane & 7;                       // 16B unit = 4 floats of the 32-dim chunk
        const int c = hh * 128 + quad * 32 + rl;      // key column of the tile
        int j = -1;
        if (c < NK) j = key0 + c;
        else if (c == NK) j = 0;
        const bool key_ok = (c < NK) ? ((kbits[c >> 5] >> (c & 31)) & 1u) != 0 : (c == NK && g_ok);
        const float4 x = *reinterpret_cast<const float4*>(slab + rl * 128 + ((u ^ (rl & 7)) << 4));
        if (key_ok) {
          float* dst = p.dkv + (static_cast<size_t>(b) * p.L + j) * 2 * E + which * E + h * AB_D + dhalf * 32 + u * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w)
                       : "memory");
        }
      }
    } else {
      // bf16 path: lane owns 8 dims (two 16-byte units) of key row rl: one 16-byte bf16x2 red.add per lane,
      // 8 key rows x 64 contiguous bytes per instruction
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        const int rl = s2 * 8 + (lane >> 2);
        const int uu = lane & 3;
        const int c = hh * 128 + quad * 32 + rl;
        int j = -1;
        if (c < NK) j = key0 + c;
        else if (c == NK) j = 0;
        const bool key_ok = (c < NK) ? ((kbits[c >> 5] >> (c & 31)) & 1u) != 0 : (c == NK && g_ok);
        const float4 x0 = *reinterpret_cast<const float4*>(slab + rl * 128 + (((2 * uu) ^ (rl & 7)) << 4));
        const float4 x1 = *reinterpret_cast<const float4*>(slab + rl * 128 + (((2 * uu + 1) ^ (rl & 7)) << 4));
        if (key_ok && c == NK) {        // the CLS key: fp32 accumulation over all tiles of the sequence
          float* dst = p.dkv_cls + ((static_cast<size_t>(b) * p.H + h) * 2 + which) * AB_D + dhalf * 32 + uu * 8;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x0.x), "f"(x0.y), "f"(x0.z), "f"(x0.w)
                       : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(x1.x), "f"(x1.y), "f"(x1.z), "f"(x1.w)
                       : "memory");
        } else if (key_ok) {
          __nv_bfloat16* dst = p.dqkv + (static_cast<size_t>(b) * p.L + j) * 3 * E + (1 + which) * E + h * AB_D + dhalf * 32 + uu * 8;
          asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(pack_bf16(x0.x, x0.y)),
                       "r"(pack_bf16(x0.z, x0.w)), "r"(pack_bf16(x1.x, x1.y)), "r"(pack_bf16(x1.z, x1.w))
                       : "memory");
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();     // TMEM, the transpose slabs, kbits and s_delta are rewritten by the next tile
  tc_fence_after();
  }   // tile loop
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// dqkv[b, 0, E:3E] = bf16(dkv_cls[b])  for sequences whose position 0 is the global token (bf16 path)
__global__ void fold_cls_kernel(const float* __restrict__ cls, const uint8_t* __restrict__ mask012, __nv_bfloat16* __restrict__ dqkv,
                                int L, int H) {
  const int b = blockIdx.x, E = H * AB_D;
  if (mask012[static_cast<size_t>(b) * L] != 2) return;
  for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) {
    const int which = i / E, hd = i % E, h = hd / AB_D, d = hd % AB_D;
    dqkv[static_cast<size_t>(b) * L * 3 * E + (1 + which) * E + hd] =
        __float2bfloat16(cls[((static_cast<size_t>(b) * H + h) * 2 + which) * AB_D + d]);
  }
}

// dqkv[:, 0:E] = bf16(dq32)   (wide windows only)
__global__ void fold_dq_kernel(const float4* __restrict__ dq, __nv_bfloat16* __restrict__ dqkv, long long T, int E) {
  const int per_row = E / 4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < T * per_row;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / per_row;
    const int c4 = static_cast<int>(idx % per_row);
    const float4 v = dq[idx];
    *reinterpret_cast<uint2*>(dqkv + t * 3 * E + c4 * 4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

// dqkv[:, E:3E] = bf16(dkv)   (fold the fp32 K/V gradient scratch into the fused gradient)
__global__ void fold_dkv_kernel(const float4* __restrict__ dkv, __nv_bfloat16* __restrict__ dqkv, long long T, int E) {
  const int per_row = 2 * E / 4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < T * per_row;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx / per_row;
    const int c4 = static_cast<int>(idx % per_row);
    const float4 v = dkv[idx];
    *reinterpret_cast<uint2*>(dqkv + t * 3 * E + E + c4 * 4) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_band_attn_bwd(const rf_attn_args* a, const void* ctx, const float* lse, const void* dctx,
                                void* dqkv, float* dkv_scratch, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && ctx && lse && dctx && dqkv && dkv_scratch, "rf_band_attn_bwd: null argument");
  RF_REQUIRE(a->D == AB_D, "rf_band_attn_bwd: head_dim %d unsupported (64 only)", a->D);
  RF_REQUIRE(a->w >= 32 && a->w % 32 == 0 && a->w <= 256,
             "rf_band_attn_bwd: one-sided window %d unsupported (multiples of 32 up to 256)", a->w);
  RF_REQUIRE(a->B > 0 && a->L >= 16 && a->H > 0, "rf_band_attn_bwd: bad shape");
  static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
  if (first_use_on_device(&attr_seen)) {
    RF_CUDA(cudaFuncSetAttribute(band_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
  }
  const int E = a->H * AB_D;
  const uint64_t L = a->L, B = a->B;
  const long long T = static_cast<long long>(B) * L;
  const CUtensorMap* tm64 = get_tmap_3d(a->qkv, B, L, 3 * E, 3 * E, L * 3 * E, 64);
  const CUtensorMap* tm16 = get_tmap_3d(a->qkv, B, L, 3 * E, 3 * E, L * 3 * E, 16);
  const CUtensorMap* tmdo = get_tmap_3d(dctx, B, L, E, E, L * E, 64);
  if (!tm64 || !tm16 || !tmdo) return RF_ERR_CUDA;
  // window segments of 65 key offsets (the forward pass may segment differently: dropout masks are absolute)
  const int nseg = (2 * a->w + 1 + 64) / 65;
  RF_REQUIRE(nseg == 1 || a->ws != nullptr, "rf_band_attn_bwd: windows wider than 64 need a workspace");
  float* dq32 = nseg > 1 ? reinterpret_cast<float*>(a->ws) : nullptr;
  const bool bf16_path = nseg == 1;
  if (bf16_path) {
    // dK/dV columns of dqkv are accumulated in place: zero them (2-D memset over columns E..3E of every row)
    RF_CUDA(cudaMemset2DAsync(reinterpret_cast<__nv_bfloat16*>(dqkv) + E, static_cast<size_t>(3) * E * 2, 0,
                              static_cast<size_t>(2) * E * 2, static_cast<size_t>(T), stream));
    RF_CUDA(cudaMemsetAsync(dkv_scratch, 0, static_cast<size_t>(B) * a->H * 2 * AB_D * sizeof(float), stream));
  } else {
    RF_CUDA(cudaMemsetAsync(dkv_scratch, 0, static_cast<size_t>(T) * 2 * E * sizeof(float), stream));
  }
  if (dq32) RF_CUDA(cudaMemsetAsync(dq32, 0, static_cast<size_t>(T) * E * sizeof(float), stream));
  AttnBwdParams p;
  p.mask012 = a->mask012; p.lse = lse;
  p.ctx = reinterpret_cast<const __nv_bfloat16*>(ctx);
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.dkv = bf16_path ? nullptr : dkv_scratch;
  p.dkv_cls = dkv_scratch;
  p.dq32 = dq32;
  p.B = a->B; p.L = a->L; p.H = a->H;
  p.drop_thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  p.drop_scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  const int tiles = (a->L + 127) / 128;
  for (int k = 0; k < nseg; ++k) {
    const int lo = -a->w + 65 * k, hi = lo + 64;
    p.shift = lo + 32;
    p.hi_cut = hi > a->w ? hi - a->w : 0;
    p.use_cls = k == 0;
    p.drop_seed = a->drop_seed;     // masks are keyed on absolute (row, key): the same seed for every segment
    const int total = a->B * a->H * tiles;
    {
      const RowActivity& ra = row_activity();
      const bool on = ra.flags != nullptr && ra.rows == static_cast<long long>(a->B) * a->L && a->L % 256 == 0;
      p.qtiles = on ? ra.qtiles : nullptr;
      p.n_qtiles = on ? ra.n_qtiles : nullptr;
    }
    band_attn_bwd_kernel<<<total < sm_count() ? total : sm_count(), AB_THREADS, AB_SMEM, stream>>>(*tm64, *tm16, *tmdo, p);
    int rc = check_launch("rf_band_attn_bwd");
    if (rc) return rc;
  }
  if (bf16_path) {
    fold_cls_kernel<<<a->B, 256, 0, stream>>>(dkv_scratch, a->mask012, reinterpret_cast<__nv_bfloat16*>(dqkv), a->L, a->H);
    return check_launch("rf_band_attn_bwd/fold_cls");
  }
  long long grid = (T * (2 * E / 4) + 255) / 256;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  fold_dkv_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(reinterpret_cast<const float4*>(dkv_scratch),
                                                             reinterpret_cast<__nv_bfloat16*>(dqkv), T, E);
  int rc = check_launch("rf_band_attn_bwd/fold");
  if (rc || !dq32) return rc;
  fold_dq_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(reinterpret_cast<const float4*>(dq32),
                                                            reinterpret_cast<__nv_bfloat16*>(dqkv), T, E);
  return check_launch("rf_band_attn_bwd/fold_dq");
}
