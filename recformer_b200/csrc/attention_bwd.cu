// Banded attention backward, sm_100a.  The kernel covers 65 key offsets per query (one-sided window 32 =
// attention_window 64); wider windows run it once per window segment (shifted keys, see attention_fwd.cu)
// with the FINAL log-sum-exp / context of the merged forward, accumulating dQ in an fp32 scratch.
//
// One work item = one (batch, head, 128-query tile), the same tiling as the forward kernel; the kernel is PERSISTENT
// (one CTA per SM walks a contiguous run of work items, see band_attn_bwd_kernel below).  Per work item:
//   S  = Q K^T,  dP = dO V^T                      tcgen05.mma -> TMEM (2 x 208 columns)
//   delta = rowsum(dO o O) (= sum_j P'_j dP_j, taken from the saved context instead of a first pass
//   over the probabilities),  P = exp(S - lse),  dS = P (dP - delta)    fp32, thread = query row
//   P, dS (bf16) -> shared memory (128B-swizzled, rows = queries)
//   dQ = dS K          A = dS (K-major),   B = K  (MN-major view of the K tile)
//   dV = P^T dO        A = P  (MN-major view of the same P buffer), B = dO (MN-major)
//   dK = dS^T Q        A = dS (MN-major view),                       B = Q  (MN-major)
// so no operand is ever transposed in memory.  dQ (scaled back by 1/sqrt(D)) is written as bf16
// into dqkv.  dK/dV tiles overlap between neighbouring query tiles: at attention_window 64 a key receives at most
// two partial sums.  Inside a CTA's run the shared keys are carried in registers to the next tile and stored once with
// plain 16-byte stores; only at the ends of a run do they go into the bf16 gradient with 16-byte
// red.add.noftz.v4.bf16x2 (the CLS key, which every tile feeds, is summed in registers per (sequence, head), then
// accumulated in fp32 and folded in by a tiny kernel); wide windows (several segments per key) accumulate dK/dV and
// dQ in fp32 scratch that a fold kernel converts.  The global query row receives no band gradient (its band output is
// overwritten by the global row, HF:615-626).
// Key validity is a bit mask per tile (band position by shifts: no per-element shared-memory flag loads).  Dropout:
// the keep bits the forward saved (rf_attn_args.keepbits, one word per thread) are read back; without them the masks
// are regenerated from ABSOLUTE (row, key) coordinates (rf_ptx.cuh), independent of the tiling — bit-identical.
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
  const uint32_t* keepbits;   // [B, H, L, 4]: the forward's dropout keep bits (attention_window 64 only: thread (row, part) reads
                              // word `part` = its 24 window columns, bit 24 of word 3 = the CLS column), or null = regenerate
  // row activity (rf_set_row_activity): the persistent CTAs walk the compact list of active query tiles x heads
  const int32_t* qtiles;      // [n] entries b * tiles_per_seq + tile, or null = every tile
  const int32_t* n_qtiles;    // device scalar n
};

#ifdef RF_KTIMING
__device__ long long g_kt_bwd[4][16][16];     // [warp 0 / 5 / 10 / 15][tile][stamp]
#define KT(k) do { if (lane == 0 && (warp % 5) == 0 && blockIdx.x == 5 && it < 16) g_kt_bwd[warp / 5][it][k] = clock64(); } while (0)
#else
#define KT(k) do {} while (0)
#endif

// PERSISTENT, one CTA per SM.  Each CTA owns a CONTIGUOUS run of the (b, h, tile) list in which the query tile is the
// fastest index, so consecutive work items are neighbouring tiles of one (sequence, head):
//  * the 64 keys that two neighbouring tiles share ([i0 + 96, i0 + 160) = tile columns 128..191 of the first = columns
//    0..63 of the second; the same TMEM lanes of the second key half / the first key half, i.e. the SAME THREADS of the
//    dK / dV epilogue) are carried in registers from one tile to the next and stored once, with plain 16-byte stores;
//    only the first / last tile of a run, whose neighbour belongs to another CTA, uses red.add on the shared keys.
//    (Before: every key of every tile went through red.global.add.v4.bf16x2 — 128 warp-level REDs per tile at
//    ~41 cycles each were 27 % of the tile time.)
//  * the CLS key's gradient is accumulated in registers over the tiles of a (sequence, head) and flushed once;
//  * the next tile's operands are TMA-loaded as soon as the current tile's last MMA has retired, its key-valid bits
//    are built from mask bytes loaded a tile ahead (no global-load latency between tiles), its context / LSE rows are
//    L2-prefetched; barriers and the 512 TMEM columns are set up once per CTA.
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
  uint32_t* kbits_all = reinterpret_cast<uint32_t*>(smem + AB_OFF_FLAG);   // [2][8]: double-buffered per tile
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + AB_OFF_BAR);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);
  uint64_t* bar_kv = bar_load + 3;     // all dK / dV accumulators ready = every MMA of the tile retired (dQ: bar_mma)
  float* s_delta = reinterpret_cast<float*>(smem + AB_OFF_DELTA);   // [4][128] partial row sums

  // 16 warps: TMEM lane quadrant = warp % 4 (rows 32*quad..), `part` = warp / 4 splits every row's work
  // between four threads (the kernel runs one CTA per SM: its warps are all the latency hiding there is):
  //   softmax backward: parts 0..2 take one 32-column window chunk each, part 3 the CLS column + the zero fill
  //   delta / dQ: 16 of the 64 head dims each;  dK / dV: part = (dK | dV, low | high 32 dims)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int tiles_per_seq = (p.L + 127) / 128;
  const int n_q = p.qtiles != nullptr ? *p.n_qtiles : 0;
  const int total_tiles = p.qtiles != nullptr ? n_q * p.H : p.B * p.H * tiles_per_seq;
  const int t_begin = static_cast<int>(static_cast<long long>(blockIdx.x) * total_tiles / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * total_tiles / gridDim.x);
  // work item t -> (sequence b, head h, query tile), the tile index fastest: dense enumeration, or head-major over the
  // list of active query tiles (which is sorted by sequence, then tile).  Positions are decoded once and then ADVANCED
  // (the divisions of a full decode per tile were ~1 500 cycles on every warp's critical path).
  struct Pos { int tile, h, b, idx; };
  auto decode = [&](int t) -> Pos {
    Pos o;
    if (p.qtiles != nullptr) {
      o.h = n_q > 0 ? t / n_q : 0;
      o.idx = t - o.h * n_q;
      const int q = p.qtiles[o.idx];
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
    if (p.qtiles != nullptr) {
      if (++o.idx == n_q) { o.idx = 0; ++o.h; }
      const int q = p.qtiles[o.idx];
      o.tile = q % tiles_per_seq;
      o.b = q / tiles_per_seq;
    } else if (++o.tile == tiles_per_seq) {
      o.tile = 0;
      if (++o.h == p.H) { o.h = 0; ++o.b; }
    }
    return o;
  };
  const int E = p.H * AB_D;
  constexpr uint32_t TM_S = 0, TM_DP = 256;                       // phase 1
  constexpr uint32_t TM_DQ = 0, TM_DV = 64, TM_DK = 192;          // phase 2 (dV: 2 x 64, dK: 2 x 64)
  const bool carry_enabled = p.dkv == nullptr;    // bf16 path (attention_window 64); wide windows accumulate in fp32 scratch

  // TMA loads of a tile's operands, in three groups (0: Q + dO and the expect_tx arrival, 1: K, 2: V; -1 = all) so
  // that three lightly loaded warps can share the ~1 100 cycles of issue; complete_tx of a group may precede the
  // expect_tx (the transaction count goes negative, the phase cannot complete before the arrival).
  auto issue_loads = [&](const Pos& q, int group) {     // one thread per group
    const int tile = q.tile, h = q.h, b = q.b;
    const int i0 = tile * 128, key0 = i0 - W + p.shift;
    if (group <= 0) {
      mbar_arrive_expect_tx(bar_load, 2 * AB_Q_BYTES + 2 * AB_KV_BYTES);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tma_load_3d(sQ + c * 8192, &tmQKV64, bar_load, h * AB_D, i0 + c * 64, b);
        tma_load_3d(sDO + c * 8192, &tmDO, bar_load, h * AB_D, i0 + c * 64, b);
      }
    }
    if (group < 0 || group == 1) {
#pragma unroll
      for (int c = 0; c < NK / 64; ++c) tma_load_3d(sK + c * 8192, &tmQKV64, bar_load, E + h * AB_D, key0 + c * 64, b);
      tma_load_3d(sK + NK * 128, &tmQKV16, bar_load, E + h * AB_D, 0, b);
    }
    if (group < 0 || group == 2) {
#pragma unroll
      for (int c = 0; c < NK / 64; ++c) tma_load_3d(sV + c * 8192, &tmQKV64, bar_load, 2 * E + h * AB_D, key0 + c * 64, b);
      tma_load_3d(sV + NK * 128, &tmQKV16, bar_load, 2 * E + h * AB_D, 0, b);
    }
  };
  // The mask byte that decides one key-valid bit of tile (b, tile): warps 0..5 = band columns, warp 6 lane 0 = the CLS
  // key.  One unconditional byte load from a clamped address (`ok` says whether it counts), so that nothing but the
  // load itself sits at the issue point: the value is consumed a tile phase later (store_kbits).
  auto kbyte_addr = [&](int b, int tile, bool& ok) -> const uint8_t* {
    const uint8_t* mrow = p.mask012 + static_cast<size_t>(b) * p.L;
    const int j = tile * 128 - W + p.shift + warp * 32 + lane;
    const bool band = warp < NK / 32;
    ok = band ? (j >= 0 && j < p.L) : (warp == NK / 32 && lane == 0);
    return mrow + ((band && ok) ? j : 0);
  };
  auto store_kbits = [&](uint32_t* kb, uint32_t kbyte, bool ok) {     // whole warps 0..6
    if (warp < NK / 32) {
      const uint32_t word = __ballot_sync(0xffffffffu, ok && kbyte == 1u);
      if (lane == 0) kb[warp] = word;
    } else if (warp == NK / 32 && lane == 0) {
      kb[7] = (p.use_cls && kbyte == 2u) ? 1u : 0u;
    }
  };

  // Per-thread global data of a tile (pulled into L2 a tile ahead, loaded at the top of the tile, first used after
  // the S / dP MMAs have been issued): two 16-byte pieces of the saved context tile, read coalesced — lane = (row
  // within a group of 4, 16-byte unit), 4 full lines per warp instruction — for delta_i = dO_i . O_i; this thread's
  // row: log-sum-exp, mask bytes and the dropout keep bits the forward saved.
  struct RowData { uint4 o[2]; float lse; uint32_t m_row, m_cls, kb; };
  const int r = quad * 32 + lane;       // query row of the tile == TMEM lane
  auto load_rows = [&](const Pos& q) -> RowData {
    RowData d;
    const int i0 = q.tile * 128;
    const uint8_t* mrow = p.mask012 + static_cast<size_t>(q.b) * p.L;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int rr = (warp * 2 + u) * 4 + (lane >> 3);
      d.o[u] = (i0 + rr < p.L) ? *reinterpret_cast<const uint4*>(p.ctx + (static_cast<size_t>(q.b) * p.L + i0 + rr) * E +
                                                                 q.h * AB_D + (lane & 7) * 8)
                               : make_uint4(0, 0, 0, 0);
    }
    const int i = i0 + r;
    const bool in_seq = i < p.L;
    const size_t rowid = (static_cast<size_t>(q.b) * p.H + q.h) * p.L + (in_seq ? i : 0);
    d.lse = in_seq ? p.lse[rowid] : 0.f;
    d.m_row = mrow[in_seq ? i : 0];
    d.m_cls = mrow[0];
    d.kb = (p.keepbits != nullptr && in_seq) ? p.keepbits[rowid * 4 + part] : 0u;
    return d;
  };

  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_kv, 1);
    fence_mbar_init();
    if (t_begin < t_end) issue_loads(decode(t_begin), -1);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  Pos cur = decode(t_begin < t_end ? t_begin : 0);
  if (t_begin < t_end && warp <= NK / 32) {
    bool ok;
    const uint8_t* ka = kbyte_addr(cur.b, cur.tile, ok);
    store_kbits(kbits_all, *ka, ok);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Control warp = warp 15 (quad 3, part 3): the lightest compute share in both phases (zero fill only in the softmax
  // backward, no second key half in the dK / dV epilogue), so the serial issue work (45 MMAs, 12 TMA loads per tile)
  // delays nobody else.
  const bool ctrl = warp == 15;
  // MMA issue: the control warp stays converged, the descriptors are warp-uniform 64-bit values (tile bases built once, the
  // k-step / chunk offsets added in the 16-byte-granular address field), only the tcgen05 instructions are
  // predicated on the elected lane
  const bool elected = elect_one();
  const uint64_t dQ_k = umma_smem_desc(smem_u32(sQ), 16, 1024), dK_k = umma_smem_desc(smem_u32(sK), 16, 1024);
  const uint64_t dDO_k = umma_smem_desc(smem_u32(sDO), 16, 1024), dV_k = umma_smem_desc(smem_u32(sV), 16, 1024);
  const uint64_t dDS_k = umma_smem_desc(smem_u32(sDS), 16, 1024);
  const uint64_t dK_mn = umma_smem_desc(smem_u32(sK), 8192, 1024), dQ_mn = umma_smem_desc(smem_u32(sQ), 8192, 1024);
  const uint64_t dDO_mn = umma_smem_desc(smem_u32(sDO), 8192, 1024);
  const uint64_t dP_mn = umma_smem_desc(smem_u32(sP), 16384, 1024), dDS_mn = umma_smem_desc(smem_u32(sDS), 16384, 1024);

  float carry[32];      // quad 0/1: the shared 64 keys' partial dK / dV of the previous tile; quad 2 (lane 0): the CLS key's sum
#pragma unroll
  for (int j = 0; j < 32; ++j) carry[j] = 0.f;
  bool carry_in = false;
  uint32_t it = 0;      // tiles processed by this CTA: parity of the once-per-tile barriers
#pragma unroll 1
  for (int t = t_begin; t < t_end; ++t, ++it) {
  KT(0);
  const int tile = cur.tile, h = cur.h, b = cur.b;
  const bool has_next = t + 1 < t_end;
  Pos nxt = cur;
  if (has_next) nxt = advance(cur);
  const int tile_n = nxt.tile, h_n = nxt.h, b_n = nxt.b;
  const bool same_bh_next = has_next && b_n == b && h_n == h;
  const bool carry_out = carry_enabled && same_bh_next && tile_n == tile + 1;
  KT(12);
  const uint32_t* kbits = kbits_all + (it & 1) * 8;
  const int i0 = tile * 128;
  const int key0 = i0 - W + p.shift;             // absolute key index of tile column 0
  const int i = i0 + r;
  const bool in_seq = i < p.L;
  RowData rd = load_rows(cur);
  uint32_t kbyte_n = 0;     // the mask byte behind one key-valid bit of the NEXT tile (turned into bits after this tile's last MMA)
  bool kbyte_ok = false;
  if (has_next && warp <= NK / 32)      // (volatile: the compiler must not sink the load down to its first use)
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(kbyte_n) : "l"(kbyte_addr(b_n, tile_n, kbyte_ok)));
  const uint64_t rowid = (static_cast<uint64_t>(b) * p.H + h) * p.L + (in_seq ? i : 0);
  const uint64_t rowbase = rowid * attn_drop_groups(p.L);
  KT(1);

  mbar_wait(bar_load, it & 1);       // every warp: dO is needed for delta below
  if (ctrl) {
    KT(2);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, NT, false, false);
    if (elected) {
#pragma unroll
      for (int k = 0; k < AB_D / 16; ++k) umma_bf16(tmem + TM_S, dQ_k + k * 2, dK_k + k * 2, idesc, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < AB_D / 16; ++k) umma_bf16(tmem + TM_DP, dDO_k + k * 2, dV_k + k * 2, idesc, k > 0 ? 1u : 0u);
      umma_commit(bar_mma);
    }
    __syncwarp();
  }
  // The registers loaded above become visible to the compiler HERE: without this fence ptxas schedules their first
  // consumers (bf16 unpacks, the lse scaling) right behind the loads, i.e. in front of the MMA issue, and the
  // global-load latency lands on the critical path of every tile.
  asm volatile("" : "+r"(rd.o[0].x), "+r"(rd.o[0].y), "+r"(rd.o[0].z), "+r"(rd.o[0].w), "+r"(rd.o[1].x), "+r"(rd.o[1].y),
                    "+r"(rd.o[1].z), "+r"(rd.o[1].w), "+f"(rd.lse), "+r"(rd.m_row), "+r"(rd.m_cls), "+r"(rd.kb));
  const float lse = rd.lse;
  const uint32_t m_row = rd.m_row, m_cls = rd.m_cls;
  const uint32_t kb_row = rd.kb;
  float keep_g = 1.f;        // dropout factor of the CLS column (part 3 uses it)
  if (part == 3 && p.drop_thresh != 0)
    keep_g = p.keepbits != nullptr ? (((kb_row >> 24) & 1u) ? p.drop_scale : 0.f)
                                   : attn_keep_cls(p.drop_seed, rowbase, p.drop_thresh, p.drop_scale);
  // ---- delta_i = sum_j P'_ij dP_ij = dO_i . O_i  (O = sum_j P'_ij V_j is the saved forward output), computed under
  //      the S / dP MMAs: 8 dims per lane (dO from the swizzled shared-memory tile), 8 lanes per row ----
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int rr = (warp * 2 + u) * 4 + (lane >> 3);
    const uint4 d = *reinterpret_cast<const uint4*>(sDO + (rr >> 6) * 8192 + (rr & 63) * 128 + (((lane & 7) ^ (rr & 7)) << 4));
    const uint4 o = rd.o[u];
    const float2 d0 = unpack_bf16(d.x), d1 = unpack_bf16(d.y), d2 = unpack_bf16(d.z), d3 = unpack_bf16(d.w);
    const float2 o0 = unpack_bf16(o.x), o1 = unpack_bf16(o.y), o2 = unpack_bf16(o.z), o3 = unpack_bf16(o.w);
    float dl = d0.x * o0.x + d0.y * o0.y + d1.x * o1.x + d1.y * o1.y + d2.x * o2.x + d2.y * o2.y + d3.x * o3.x + d3.y * o3.y;
    dl += __shfl_xor_sync(0xffffffffu, dl, 1);
    dl += __shfl_xor_sync(0xffffffffu, dl, 2);
    dl += __shfl_xor_sync(0xffffffffu, dl, 4);
    if ((lane & 7) == 0) s_delta[rr] = dl;
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  KT(3);

  const bool is_global_row = (i == 0) && (m_cls == 2);
  const bool row_valid = in_seq && (m_row != 0) && !is_global_row;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const float LOG2E = 1.4426950408889634f;
  const float lse2 = lse * LOG2E;
  const bool g_ok = kbits[7] != 0;
  const int band_hi = 2 * W - p.hi_cut;      // a row's live columns: key-valid bits AND band position [r, r + band_hi]

  float pg = 0.f, pg_d = 0.f, dpg = 0.f;
  if (part == 3) {   // warp-uniform: the CLS column
    uint32_t gs[8], gd[8];
    tmem_ld8(lane_base + TM_S + NK, gs);
    tmem_ld8(lane_base + TM_DP + NK, gd);
    tmem_ld_wait();
    pg = (row_valid && g_ok) ? exp2f(__uint_as_float(gs[0]) * LOG2E - lse2) : 0.f;   // undropped
    pg_d = pg * keep_g;
    dpg = __uint_as_float(gd[0]);
  }
  __syncthreads();
  KT(4);
  // masked rows may hold non-finite garbage in O / dO
  const float delta = row_valid ? s_delta[r] : 0.f;

  // ---- pass B: P' and dS -> shared memory.  The 96 window columns [32 quad, 32 quad + 96) of a row quadrant are split
  //      evenly between the four parts: 24 columns (three 8-key groups) each ----
  {
    const int c0 = quad * 32 + part * 24;                 // first tile column of this thread's piece
    uint32_t sv[24], dv[24];
    {
      uint32_t a16[16], a8[8], b16[16], b8[8];
      tmem_ld16(lane_base + TM_S + c0, a16);
      tmem_ld8(lane_base + TM_S + c0 + 16, a8);
      tmem_ld16(lane_base + TM_DP + c0, b16);
      tmem_ld8(lane_base + TM_DP + c0 + 16, b8);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) { sv[j] = a16[j]; dv[j] = b16[j]; }
#pragma unroll
      for (int j = 0; j < 8; ++j) { sv[16 + j] = a8[j]; dv[16 + j] = b8[j]; }
    }
    uint32_t live;
    {
      const int wi = c0 >> 5, sh = c0 & 31;
      const uint64_t kw = (static_cast<uint64_t>(kbits[wi + 1]) << 32) | kbits[wi];     // wi + 1 <= 6 (unused word: masked off)
      const int lo = r - c0, hi = r + band_hi - c0;
      const uint32_t mlo = lo <= 0 ? 0xFFFFFFu : (lo >= 24 ? 0u : ((0xFFFFFFu << lo) & 0xFFFFFFu));
      const uint32_t mhi = hi >= 23 ? 0xFFFFFFu : (hi < 0 ? 0u : (0xFFFFFFu >> (23 - hi)));
      live = row_valid ? (static_cast<uint32_t>(kw >> sh) & mlo & mhi) : 0u;
    }
    uint32_t keepm = 0xFFFFFFu;
    if (p.drop_thresh != 0) {
      if (p.keepbits != nullptr) {
        keepm = kb_row;       // the forward saved this thread's 24 keep bits
      } else if (live != 0) {
        keepm = attn_keep32(p.drop_seed, rowbase, key0 + c0, p.drop_thresh, live);
      }
    }
    uint32_t po[12], so[12];
#pragma unroll
    for (int j = 0; j < 24; j += 2) {
      float pr[2], ds[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pu = ((live >> (j + e)) & 1u) ? exp2f(__uint_as_float(sv[j + e]) * LOG2E - lse2) : 0.f;
        const float kp = ((keepm >> (j + e)) & 1u) ? p.drop_scale : 0.f;
        pr[e] = pu * kp;                                               // P' feeds dV
        ds[e] = pu * (kp * __uint_as_float(dv[j + e]) - delta);        // softmax backward
      }
      po[j >> 1] = pack_bf16(pr[0], pr[1]);
      so[j >> 1] = pack_bf16(ds[0], ds[1]);
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int c = c0 + u * 8;
      const uint32_t o = (c >> 6) * 16384 + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(sP + o) = make_uint4(po[u * 4], po[u * 4 + 1], po[u * 4 + 2], po[u * 4 + 3]);
      *reinterpret_cast<uint4*>(sDS + o) = make_uint4(so[u * 4], so[u * 4 + 1], so[u * 4 + 2], so[u * 4 + 3]);
    }
    // Zero fill of the 96 band columns outside the window (twelve 8-column units, three per part) and of the global
    // chunk = P/dS chunk 3 (columns 192..255: column 192 holds the CLS key, the rest is zero).  A thread writes the same
    // window columns of its row every tile, so in the bf16 path — where nothing else ever touches the P / dS buffers —
    // the fill is needed for the CTA's first tile only; the fp32 path reuses the P buffer as transpose slabs.
    if (!carry_enabled || it == 0) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int k = part * 3 + u;
        const int cu = k < 4 * quad ? k : k + 12;           // 8-column unit of the tile
        const uint32_t o = (cu >> 3) * 16384 + r * 128 + (((cu & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sP + o) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sDS + o) = make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {      // units 1..7 of the global chunk: part 3 takes 1, parts 0..2 units 2..7
        const int u = ((part + 1) & 3) * 2 + uu;
        if (u == 0) continue;
        const uint32_t o = 3 * 16384 + r * 128 + ((u ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sP + o) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sDS + o) = make_uint4(0, 0, 0, 0);
      }
    }
    if (part == 3) {     // the CLS column itself (unit 0 of the global chunk), every tile
      const float dsg = pg * (keep_g * dpg - delta);
      const uint32_t o = 3 * 16384 + r * 128 + ((r & 7) << 4);
      *reinterpret_cast<uint4*>(sP + o) = make_uint4(pack_bf16(pg_d, 0.f), 0, 0, 0);
      *reinterpret_cast<uint4*>(sDS + o) = make_uint4(pack_bf16(dsg, 0.f), 0, 0, 0);
    }
  }

  KT(5);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  KT(6);
  if (ctrl) {
    tc_fence_after();
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, AB_D, false, true);
    constexpr uint32_t idesc_kv = umma_idesc_bf16(128, AB_D, true, true);
    if (elected) {
      // dQ[128 x 64] = dS[128 x 208] K[208 x 64]   (dS chunk of 64 keys = +16384 B, k-step = +32 B; K rows: +2048 B)
#pragma unroll
      for (int ks = 0; ks < NT / 16; ++ks)
        umma_bf16(tmem + TM_DQ, dDS_k + ((ks >> 2) * 1024 + (ks & 3) * 2), dK_mn + ks * 128, idesc_q, ks > 0 ? 1u : 0u);
      umma_commit(bar_mma);   // dQ is stored while the dK / dV MMAs below still run
      // dV[keys x 64] = P^T dO,  dK[keys x 64] = dS^T Q   (two 128-key halves each)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int ks = 0; ks < 128 / 16; ++ks) {
          umma_bf16(tmem + TM_DV + hh * 64, dP_mn + (hh * 2048 + ks * 128), dDO_mn + ks * 128, idesc_kv, ks > 0 ? 1u : 0u);
          umma_bf16(tmem + TM_DK + hh * 64, dDS_mn + (hh * 2048 + ks * 128), dQ_mn + ks * 128, idesc_kv, ks > 0 ? 1u : 0u);
        }
      }
      umma_commit(bar_kv);
    }
    __syncwarp();
  }
  mbar_wait(bar_mma, 1);
  tc_fence_after();
  KT(7);

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
      // bf16 path: this thread's 16 dims are one full 32-byte sector of the row (the four parts complete the 128-byte
      // line): a single 256-bit store
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = pack_bf16(__uint_as_float(v[2 * j]) * 0.125f, __uint_as_float(v[2 * j + 1]) * 0.125f);
      st_global_v8(p.dqkv + (static_cast<size_t>(b) * p.L + i) * 3 * E + h * AB_D + part * 16, o);
    }
  }
  KT(8);
  // Once every MMA of this tile has retired, Q / dO / K / V are dead: the next tile's loads are issued (lane 0 of
  // warps 3, 7, 11: quadrant 3 has no second key half to store), its key-valid bits built from the bytes loaded at the
  // top of this tile, and its per-row global data (saved context rows, log-sum-exps, keep bits) pulled into L2.
  auto next_tile_setup = [&]() {
    if (!has_next) return;
    if (quad == 3 && part < 3 && lane == 0) issue_loads(nxt, part);
    __syncwarp();
    asm volatile("" : "+r"(kbyte_n));      // (first use of the byte loaded at the top of the tile: see the fence above)
    if (warp <= NK / 32) store_kbits(kbits_all + ((it + 1) & 1) * 8, kbyte_n, kbyte_ok);
    const size_t row0 = (static_cast<size_t>(b_n) * p.H + h_n) * p.L + tile_n * 128;
    if (tid < 128) {
      if (tile_n * 128 + tid < p.L)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.ctx + (static_cast<size_t>(b_n) * p.L + tile_n * 128 + tid) * E + h_n * AB_D));
    } else if (tid < 132) {
      if (tile_n * 128 + (tid - 128) * 32 < p.L) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.lse + row0 + (tid - 128) * 32));
    } else if (tid < 148 && p.keepbits != nullptr) {
      if (tile_n * 128 + (tid - 132) * 8 < p.L) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.keepbits + (row0 + (tid - 132) * 8) * 4));
    }
  };
  // (Publishing the first key half's accumulators early and draining them under the second half's MMAs was measured
  // slower: the tcgen05.ld traffic delays the MMAs, which are shared-memory-bandwidth bound at N = 64.)
  mbar_wait(bar_kv, it & 1);
  tc_fence_after();
  KT(9);
  next_tile_setup();
  KT(13);
  // every shared-memory operand is dead now: the P region becomes 16 per-warp 4 KB transpose slabs
  uint8_t* slab = sP + warp * 4096;
  // ---- dK / dV: TMEM lane = key column c = hh*128 + r of the tile.  Each 32-key x 32-dim chunk is
  //      bf16 path: straight from registers (below); fp32 path: through the warp's transpose slab. ----
  const int which = part >> 1, dhalf = part & 1;   // this warp: dK (0) or dV (1), dims 32*dhalf .. +31
#pragma unroll 1
  for (int hh = 0; hh < 2; ++hh) {    // the two 128-key halves of the tile
    if (carry_enabled && hh == 1 && quad == 3) continue;          // tile columns 224..255: nothing there
    const uint32_t tcol = (which == 0 ? TM_DK : TM_DV) + hh * 64 + dhalf * 32;
    uint32_t v[32];
    tmem_ld32(lane_base + tcol, v);
    tmem_ld_wait();
    bool use_red = true;          // warp-uniform: may another CTA (or a non-adjacent tile) add to these keys?
    if (carry_enabled) {
      if (hh == 0) {
        if (quad < 2) {           // keys shared with the previous tile
          if (carry_in) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + carry[j]);
            use_red = false;
          }
        } else {
          use_red = false;        // keys [i0 + 32, i0 + 96): only this tile's queries see them
        }
      } else if (quad < 2) {      // keys shared with the next tile
        if (carry_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) carry[j] = __uint_as_float(v[j]);
          continue;               // stored by the next tile
        }
      } else {                    // quad 2: lane 0 = the CLS key column (the other lanes hold zeros)
#pragma unroll
        for (int j = 0; j < 32; ++j) carry[j] += __uint_as_float(v[j]);
        if (!same_bh_next) {      // last tile of this (sequence, head) in this CTA's run: flush, fp32
          if (lane == 0 && p.use_cls) {
            float* dst = p.dkv_cls + ((static_cast<size_t>(b) * p.H + h) * 2 + which) * AB_D + dhalf * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(carry[j]), "f"(carry[j + 1]),
                           "f"(carry[j + 2]), "f"(carry[j + 3])
                           : "memory");
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) carry[j] = 0.f;
        }
        continue;
      }
    }
    if (p.dkv == nullptr) {
      // bf16 path (band keys only: the CLS column went into `carry` above): this thread = key row c, its 32 dims are
      // 64 contiguous bytes of the gradient row: two 256-bit stores, or four bf16x2 red.adds on keys another CTA shares
      const int c = hh * 128 + quad * 32 + lane;      // < NK here
      if ((kbits[c >> 5] >> (c & 31)) & 1u) {
        __nv_bfloat16* dst = p.dqkv + (static_cast<size_t>(b) * p.L + key0 + c) * 3 * E + (1 + which) * E + h * AB_D + dhalf * 32;
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        if (use_red) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst + 2 * j), "r"(o[j]), "r"(o[j + 1]),
                         "r"(o[j + 2]), "r"(o[j + 3])
                         : "memory");
        } else {
          st_global_v8(dst, o);
          st_global_v8(dst + 16, o + 8);
        }
      }
      continue;
    }
    // fp32 path (wide windows): the 32-key x 32-dim chunk is transposed through the warp's slab so that one red.add.v4
    // instruction covers 4 key rows x 128 contiguous bytes (4 LSU wavefronts) instead of 32 rows x 16 bytes
    {
      uint8_t* srow = slab + lane * 128;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<uint4*>(srow + ((u ^ (lane & 7)) << 4)) = make_uint4(v[u * 4], v[u * 4 + 1], v[u * 4 + 2], v[u * 4 + 3]);
    }
    __syncwarp();
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
    __syncwarp();
  }
  carry_in = carry_out;
  cur = nxt;

  KT(10);
  tc_fence_before();
  __syncthreads();     // TMEM, the transpose slabs, the other kbits buffer and s_delta are rewritten by the next tile
  tc_fence_after();
  KT(11);
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

#ifdef RF_KTIMING
extern "C" int rf_debug_ktiming_bwd(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_kt_bwd, sizeof(g_kt_bwd)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int rf_band_attn_bwd(const rf_attn_args* a, const void* ctx, const float* lse, const void* dctx,
                                void* dqkv, float* dkv_scratch, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && ctx && lse && dctx && dqkv && dkv_scratch, "rf_band_attn_bwd: null argument");
  RF_REQUIRE(a->D == AB_D, "rf_band_attn_bwd: head_dim %d unsupported (64 only)", a->D);
  RF_REQUIRE(a->w >= 32 && a->w % 32 == 0 && a->w <= 256,
             "rf_band_attn_bwd: one-sided window %d unsupported (multiples of 32 up to 256)", a->w);
  RF_REQUIRE(a->B > 0 && a->L >= 16 && a->H > 0, "rf_band_attn_bwd: bad shape");
  RF_REQUIRE((reinterpret_cast<uintptr_t>(dqkv) & 31) == 0, "rf_band_attn_bwd: dqkv must be 32-byte aligned (256-bit stores)");
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
  p.keepbits = (bf16_path && a->drop_p > 0.f) ? reinterpret_cast<const uint32_t*>(a->keepbits) : nullptr;
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
