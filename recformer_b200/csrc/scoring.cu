// Full-catalogue cosine scoring (SURVEY.md §8a Spec S; ref: recformer/models.py:358-369,539-545):
//   logit[b,n] = (x_b / |x_b|) . (y_n / |y_n|) / temp
// as a bf16 tcgen05 GEMM over a pre-normalised item table with fused epilogues:
//   TOPK : temperature scaling + streaming per-row top-k, logits never written.  Each epilogue
//          thread owns one user row x one 128-column half of the tile; its sorted k-list lives in
//          shared memory and only the k-th best score is kept in a register, so a 32-column
//          accumulator chunk that cannot enter the list costs one max per column + one warp vote;
//          the label score is picked out of the same accumulator so that Spec R's strict `>`
//          rank count is exact;
//   DENSE: fp32 logits [B,N] (the (B,N) score tensor RecformerForSeqRec.forward returns, and the
//          training cross-entropy input, ref: recformer/models.py:583-591).
// Scheduling: CTA = (user m-tile, item slice s); slice s sweeps item tiles s, s+S, ... so all
// m-tiles advance through the table in lock-step and each 256-item tile is fetched from HBM once
// and re-read from L2 by the other m-tiles.  More than 128 users: CTA pairs (cta_group::2) on
// 256 x 256 tiles; otherwise single CTAs on 128 x 256 tiles.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

namespace rf {

constexpr int SC_BM = 128, SC_BN = 256, SC_BK = 64, SC_STAGES = 3;
constexpr int SC_EPI_WARPS = 8, SC_THREADS = 64 + 32 * SC_EPI_WARPS;
constexpr int SC_MAXK = 16;
constexpr uint32_t SC_A_BYTES = SC_BM * SC_BK * 2, SC_B_BYTES = SC_BN * SC_BK * 2;
constexpr uint32_t SC_STAGE_BYTES = SC_A_BYTES + SC_B_BYTES;
constexpr uint32_t SC_LIST_ONLY_BYTES = SC_MAXK * SC_EPI_WARPS * 32 * 8;   // per-row top-k lists: scores, then ids
constexpr uint32_t SC_LIST_BYTES = SC_LIST_ONLY_BYTES + SC_EPI_WARPS * 4096;   // + per-warp [32 cols][32 rows] fp32 chunk slab
constexpr uint32_t SC_SMEM = SC_STAGES * SC_STAGE_BYTES + SC_LIST_BYTES + (2 * SC_STAGES + 4) * 8 + 16 + 1024;
// CTA-pair variant: 256 users x 256 items per tile, each CTA stages its 128 user rows and its
// 128-item half of the table slab (32 KB per 64-wide K slab), 5 ring stages
constexpr int SP_STAGES = 5;
constexpr uint32_t SP_STAGE_BYTES = 2 * 128 * SC_BK * 2;
constexpr uint32_t SP_SMEM = SP_STAGES * SP_STAGE_BYTES + SC_LIST_BYTES + (2 * SP_STAGES + 4) * 8 + 16 + 1024;

struct ScoreParams {
  int B; long long N; int K;      // users, items, hidden
  float inv_temp;
  int k;                          // top-k
  int id_base;
  int slices;                     // S: item-tile slices per user tile (lock-step CTAs / pairs)
  int extra;                      // pair kernel: leftover pairs (SM pairs not divisible by the user tiles)
  int main_tiles;                 // item tiles [0, main_tiles) belong to the lock-step pairs
  int extra_tiles;                // item tiles [main_tiles, main_tiles + extra_tiles) belong to the leftover pairs
  int extra_steps;                // (user tile, item tile) steps per leftover pair
  const int64_t* labels;
  float* ws_scores;               // [parts][B][k]   part = (slice or S + extra pair) * 2 + column half
  int32_t* ws_ids;                // [parts][B][k]
  float* ws_label;                // [parts][B]
  unsigned int* ws_thr;           // [B + 1]: per user row, max over all parts of their current k-th best score (order-
                                  //          preserving uint encoding, 0 = nothing yet); slot B absorbs rows past B
  float* logits;                  // DENSE: [B][N]
};

enum { SC_TOPK = 0, SC_DENSE = 1 };

// Per-thread (= per user row) streaming top-k state.  The sorted list lives in shared memory
// ([q][thread], conflict-free); only the current k-th best score (`thr`) is kept in a register, so
// the common case — no column of a 32-column accumulator chunk beats thr — costs one max per
// column and a single warp vote.
//
// Cross-part pruning.  A user row's items are swept by several parts (item-tile slices x column halves, leftover
// pairs), each with its own list.  A part whose list is full publishes its k-th best score t (atomicMax on an order-
// preserving encoding): that part alone holds k items scoring >= t, so no item scoring < t can be in the row's global
// top-k.  Every part therefore also rejects candidates below the published maximum g — with `>=`, not `>`: an item
// that TIES with g may still win its place by the lower-id rule.  Without this every part pays the full warm-up of a
// young list (k (1 + ln(n/k)) inserts for its n items); with it the row pays roughly one warm-up in total, which is
// what makes a 125k-item shard (8-GPU sharding of the 1M table) cost little more per item than the whole table.
//
// List layout and the two insert paths.  A warp's 32 rows keep their sorted lists in one shared-memory block, entry q of
// row t at word t * 16 + (q ^ (t >> 1)): conflict-free both when every lane touches entry q of its OWN row (banks
// (t & 1) * 16 + (q ^ (t >> 1)) are all different) and when lanes 0..k-1 touch the k entries of ONE row.  A chunk that
// produced hits inserts them either
//   * per lane (SIMT): every lane with hits runs a compare-and-shift insertion on its own row — the warp pays
//     max-over-lanes(hits) x ~90 instructions of a dependent shared-memory chain, or
//   * cooperatively: the warp takes the hits one by one, lane q holds entry q of the hit's row, one ballot gives the
//     insertion position, one shuffle shifts the tail (~20 instructions per hit, total-over-lanes of them).
// The chunk picks whichever is cheaper from the hit counts (one redux each).  Young lists (first tiles of a part, or
// the short item range a shard of an 8-GPU run sees) have many hits per lane: SIMT; afterwards a chunk typically has
// a handful of hits spread over different lanes, where the SIMT path idles 31 lanes per insertion: cooperative.
struct TopkState {
  float* slab;      // this warp's [32 cols][32 rows] fp32 chunk slab
  float* ls;        // this warp's list scores [32 rows][16 entries] (swizzled, see list_slot)
  int* li;          // ... and ids
  float thr;        // k-th best score of THIS part's list (-inf until the list is full)
  float eff;        // strict rejection threshold: max(thr, largest float below the published g)
  float gprev;      // largest float below g
  float published;  // last value of thr this thread published
  unsigned int* gthr;
  float label_score;
  long long label_local;   // label id relative to this table shard, or -1
};

__device__ __forceinline__ int list_slot(int t, int q) { return t * SC_MAXK + (q ^ (t >> 1)); }

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
constexpr unsigned int ORD_NEG_INF = 0x007FFFFFu;   // f2ord(-inf); anything <= is "no threshold yet"

// once per item tile: pick up what the other parts have published, publish our own k-th best if it rose
__device__ __forceinline__ void topk_sync_threshold(TopkState& st) {
  if (st.thr > st.published) {
    atomicMax(st.gthr, f2ord(st.thr));
    st.published = st.thr;
  }
  const unsigned int u = *reinterpret_cast<volatile unsigned int*>(st.gthr);
  st.gprev = (u <= ORD_NEG_INF) ? -INFINITY : ord2f(u - 1);
  st.eff = fmaxf(st.thr, st.gprev);
}

// per-lane insertion into the lane's own row; returns the new k-th best
__device__ __forceinline__ float topk_insert_own(const TopkState& st, int lane, int k, float s, int id) {
  // descending insertion; strict '>' keeps the earlier (lower) id ahead on equal scores
  int q = k - 1;
  while (q > 0) {
    const float prev = st.ls[list_slot(lane, q - 1)];
    if (!(s > prev)) break;
    st.ls[list_slot(lane, q)] = prev;
    st.li[list_slot(lane, q)] = st.li[list_slot(lane, q - 1)];
    --q;
  }
  st.ls[list_slot(lane, q)] = s;
  st.li[list_slot(lane, q)] = id;
  return st.ls[list_slot(lane, k - 1)];
}

// warp-cooperative insertion of (s, id) into row L's list (all arguments warp-uniform); returns the new k-th best
__device__ __forceinline__ float topk_insert_coop(const TopkState& st, int lane, int k, int L, float s, int id) {
  float e_s = -INFINITY;
  int e_i = 0x7fffffff;
  if (lane < SC_MAXK) { e_s = st.ls[list_slot(L, lane)]; e_i = st.li[list_slot(L, lane)]; }
  const bool ahead = lane < k && (e_s > s || (e_s == s && e_i < id));     // a prefix of the sorted list
  const int pos = __popc(__ballot_sync(0xffffffffu, ahead));
  const float up_s = __shfl_up_sync(0xffffffffu, e_s, 1);
  const int up_i = __shfl_up_sync(0xffffffffu, e_i, 1);
  float n_s = e_s;
  int n_i = e_i;
  if (lane == pos) { n_s = s; n_i = id; }
  else if (lane > pos) { n_s = up_s; n_i = up_i; }
  if (lane >= pos && lane < k) { st.ls[list_slot(L, lane)] = n_s; st.li[list_slot(L, lane)] = n_i; }
  const float thr_new = __shfl_sync(0xffffffffu, n_s, k - 1);
  __syncwarp();
  return thr_new;
}

// One 32-row x 32-column accumulator chunk (thread = user row) of the TOPK epilogue.
__device__ __forceinline__ void topk_chunk(const uint32_t (&r)[32], TopkState& st, const ScoreParams& p,
                                           long long col0) {
  const int lane = threadIdx.x & 31;
  // fast reject: max over the chunk (4 independent chains), one multiply, one vote
  float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
  for (int j = 4; j < 32; j += 4) {
    m0 = fmaxf(m0, __uint_as_float(r[j]));
    m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
    m2 = fmaxf(m2, __uint_as_float(r[j + 2]));
    m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
  }
  const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * p.inv_temp;   // == max_j (r[j] * inv_temp): rounding is monotonic
  const long long rel = st.label_local - col0;
  const bool partial = col0 + 32 > p.N;                                // warp-uniform
  const int nvalid = partial ? static_cast<int>(p.N - col0) : 32;
  const bool has_label = rel >= 0 && rel < nvalid;
  const bool mine = (mx > st.eff) || has_label;
  if (!__any_sync(0xffffffffu, mine)) return;
  // slow path.  Bit mask of the columns that MAY beat the threshold: the raw accumulator against a threshold lowered
  // by a few ulps (one compare per column, no multiply); the exact test s * inv_temp > eff is repeated on each
  // candidate at insertion time — the threshold may have risen by then anyway.
  uint32_t hits = 0;
  if (mx > st.eff) {
    float raw = __fdividef(st.eff, p.inv_temp);
    raw = raw - fabsf(raw) * 4e-6f - 1e-30f;          // (-inf stays -inf)
#pragma unroll
    for (int j = 0; j < 32; ++j) hits |= (__uint_as_float(r[j]) > raw) ? (1u << j) : 0u;
    if (partial) hits &= (1u << nvalid) - 1u;
  }
  // the chunk's values go through the warp's slab so that a hit (or the label column) can be fetched by its
  // dynamic column index, by any lane
#pragma unroll
  for (int j = 0; j < 32; ++j) st.slab[j * 32 + lane] = __uint_as_float(r[j]);
  __syncwarp();
  if (has_label) st.label_score = st.slab[static_cast<int>(rel) * 32 + lane] * p.inv_temp;
  const int nh = __popc(hits);
  const int total = __reduce_add_sync(0xffffffffu, nh);
  const int id0 = p.id_base + static_cast<int>(col0);
  if (total != 0) {
    const int most = __reduce_max_sync(0xffffffffu, nh);
    if (total > 4 * most) {
      // ---- per-lane insertions ----
      while (hits != 0) {
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        const float s = st.slab[j * 32 + lane] * p.inv_temp;
        if (s > st.eff) {
          st.thr = topk_insert_own(st, lane, p.k, s, id0 + j);
          st.eff = fmaxf(st.thr, st.gprev);
        }
      }
    } else {
      // ---- cooperative insertions, hit by hit ----
      uint32_t lanes = __ballot_sync(0xffffffffu, hits != 0);
      while (lanes != 0) {
        const int L = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        uint32_t m = __shfl_sync(0xffffffffu, hits, L);
        while (m != 0) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          const float s = st.slab[j * 32 + L] * p.inv_temp;
          const float eff_l = __shfl_sync(0xffffffffu, st.eff, L);
          if (s > eff_l) {
            const float thr_new = topk_insert_coop(st, lane, p.k, L, s, id0 + j);
            if (lane == L) { st.thr = thr_new; st.eff = fmaxf(thr_new, st.gprev); }
          }
        }
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void dense_chunk(const uint32_t (&r)[32], const ScoreParams& p, int row, bool row_ok,
                                            long long col0) {
  if (!row_ok) return;
  float* out = p.logits + static_cast<size_t>(row) * p.N + col0;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (col0 + j < p.N) out[j] = __uint_as_float(r[j]) * p.inv_temp;
}

__device__ __forceinline__ void topk_state_init(TopkState& st, uint8_t* list_smem, int etid, const ScoreParams& p, int row,
                                                bool row_ok) {
  const int ewarp = etid >> 5, lane = etid & 31;
  st.slab = reinterpret_cast<float*>(list_smem + SC_LIST_ONLY_BYTES) + ewarp * 1024;
  st.ls = reinterpret_cast<float*>(list_smem) + ewarp * 32 * SC_MAXK;
  st.li = reinterpret_cast<int*>(list_smem + SC_EPI_WARPS * 32 * SC_MAXK * 4) + ewarp * 32 * SC_MAXK;
  __syncwarp();     // (re-initialisation: cooperative reads of the previous segment's lists are over)
  for (int q = 0; q < SC_MAXK; ++q) { st.ls[list_slot(lane, q)] = -INFINITY; st.li[list_slot(lane, q)] = 0x7fffffff; }
  __syncwarp();
  st.thr = st.eff = st.gprev = st.published = -INFINITY;
  st.gthr = p.ws_thr + (row_ok ? row : p.B);
  st.label_score = -INFINITY;
  st.label_local = -1;
  if (p.labels != nullptr && row_ok) {
    const long long l = p.labels[row] - p.id_base;
    st.label_local = (l >= 0 && l < p.N) ? l : -1;
  }
}

__device__ __forceinline__ void topk_state_flush(const TopkState& st, const ScoreParams& p, int part, int row) {
  const int lane = threadIdx.x & 31;
  float* os = p.ws_scores + (static_cast<size_t>(part) * p.B + row) * p.k;
  int32_t* oi = p.ws_ids + (static_cast<size_t>(part) * p.B + row) * p.k;
  for (int q = 0; q < p.k; ++q) { os[q] = st.ls[list_slot(lane, q)]; oi[q] = st.li[list_slot(lane, q)]; }
  p.ws_label[static_cast<size_t>(part) * p.B + row] = st.label_score;
}

// ---- single-CTA kernel (<= 128 users per tile): small user batches ----
template <int MODE>
__global__ void __launch_bounds__(SC_THREADS, 1)
cosine_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ScoreParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_list = smem + SC_STAGES * SC_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_list + SC_LIST_BYTES);
  uint64_t* empty_bar = full_bar + SC_STAGES;
  uint64_t* tfull_bar = empty_bar + SC_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / p.slices;
  const int slice = blockIdx.x % p.slices;
  const int m0 = mt * SC_BM;
  const int n_tiles = static_cast<int>((p.N + SC_BN - 1) / SC_BN);
  const int k_blocks = (p.K + SC_BK - 1) / SC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], SC_EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * SC_BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0; uint32_t phase = 0;
      for (int nt = slice; nt < n_tiles; nt += p.slices) {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], SC_STAGE_BYTES);
          uint8_t* sa = smem + stage * SC_STAGE_BYTES;
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * SC_BK, m0);
          tma_load_2d(sa + SC_A_BYTES, &tmB, &full_bar[stage], kb * SC_BK, nt * SC_BN);
          if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(SC_BM, SC_BN, false, false);
    const bool elected = elect_one();
    const uint64_t adesc0 = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t bdesc0 = umma_smem_desc(smem_u32(smem) + SC_A_BYTES, 16, 1024);
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int nt = slice; nt < n_tiles; nt += p.slices) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t soff = static_cast<uint32_t>(stage) * (SC_STAGE_BYTES >> 4);
        if (elected) {
#pragma unroll
          for (int k = 0; k < SC_BK / 16; ++k)
            umma_bf16(tmem_base + acc * SC_BN, adesc0 + soff + k * 2, bdesc0 + soff + k * 2, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // epilogue warps 2..9: TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = m0 + quad * 32 + lane;
    const bool row_ok = row < p.B;
    TopkState st;
    if (MODE == SC_TOPK) topk_state_init(st, s_list, threadIdx.x - 64, p, row, row_ok);
    int acc = 0; uint32_t acc_phase = 0;
    for (int nt = slice; nt < n_tiles; nt += p.slices) {
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const long long n0 = static_cast<long long>(nt) * SC_BN + half * (SC_BN / 2);
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * SC_BN + half * (SC_BN / 2);
      uint32_t rbuf[2][32];
      tmem_ld32(tbase, rbuf[0]);
      if (MODE == SC_TOPK) topk_sync_threshold(st);
#pragma unroll
      for (int c = 0; c < SC_BN / 2 / 32; ++c) {
        tmem_ld_wait();
        if (c + 1 < SC_BN / 2 / 32) {
          tmem_ld32(tbase + (c + 1) * 32, rbuf[(c + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        const long long col0 = n0 + c * 32;
        if (col0 >= p.N) continue;
        if (MODE == SC_DENSE) dense_chunk(rbuf[c & 1], p, row, row_ok, col0);
        else topk_chunk(rbuf[c & 1], st, p, col0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (MODE == SC_TOPK && row_ok) topk_state_flush(st, p, slice * 2 + half, row);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * SC_BN); }
}

// ---- CTA-pair kernel: 256 users x 256 items per tile with tcgen05.mma.cta_group::2 ----
// Same protocol as gemm_pair_kernel (gemm.cu): both CTAs stream their halves of the operands, the
// leader's MMA thread issues for both, commits are multicast; each CTA's 8 epilogue warps scan its own
// 128 accumulator rows.  L2 -> SM traffic per FLOP is 2/3 of the single-CTA kernel's.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SC_THREADS, 1)
cosine_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const ScoreParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_list = smem + SP_STAGES * SP_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_list + SC_LIST_BYTES);
  uint64_t* empty_bar = full_bar + SP_STAGES;
  uint64_t* tfull_bar = empty_bar + SP_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int m_tiles = (p.B + 255) / 256;
  const int n_tiles = static_cast<int>((p.N + SC_BN - 1) / SC_BN);
  const int k_blocks = (p.K + SC_BK - 1) / SC_BK;
  // Work of this pair = a list of segments (user tile, item-tile range).  Lock-step pairs own one user
  // tile and every S-th item tile of [0, main_tiles): all user tiles advance through the table together,
  // so a table tile is fetched from HBM once and re-read from L2.  The leftover pairs (SM pairs not
  // divisible by the user tiles) split the (user tile x item tile) rectangle over the remaining item
  // tiles into equal contiguous runs in user-tile-major order: few, long segments per pair (every new
  // segment restarts its top-k list, and a young list inserts often).
  const int n_main = m_tiles * p.slices;
  const bool is_extra = pair >= n_main;
  const int part = is_extra ? p.slices + (pair - n_main) : pair % p.slices;
  const long long f_lo = is_extra ? static_cast<long long>(pair - n_main) * p.extra_steps : 0;
  const long long f_end = static_cast<long long>(p.extra_tiles) * m_tiles;
  const long long f_hi = is_extra ? (f_lo + p.extra_steps < f_end ? f_lo + p.extra_steps : f_end) : 0;
  struct Seg { int mt, lo, hi, step; };
  auto first_seg = [&](long long& f) -> bool { f = f_lo; return is_extra ? f < f_hi : true; };
  auto get_seg = [&](long long f) -> Seg {
    if (!is_extra) return Seg{pair / p.slices, pair % p.slices, p.main_tiles, p.slices};
    const int mt = static_cast<int>(f / p.extra_tiles), r0 = static_cast<int>(f % p.extra_tiles);
    const long long left = f_hi - f;
    const int r1 = (p.extra_tiles - r0 < left) ? p.extra_tiles : r0 + static_cast<int>(left);
    return Seg{mt, p.main_tiles + r0, p.main_tiles + r1, 1};
  };
  auto next_seg = [&](long long& f, const Seg& sg) -> bool {
    if (!is_extra) return false;
    f += sg.hi - sg.lo;
    return f < f_hi;
  };
  const long long my_tiles = is_extra ? (f_hi > f_lo ? f_hi - f_lo : 0)
                                      : (p.main_tiles > pair % p.slices
                                             ? (p.main_tiles - pair % p.slices + p.slices - 1) / p.slices : 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < SP_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * SC_EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0; uint32_t phase = 0;
      long long f;
      for (bool more = first_seg(f); more;) {
        const Seg sg = get_seg(f);
        const int m0 = sg.mt * 256 + static_cast<int>(rank) * 128;
        for (int nt = sg.lo; nt < sg.hi; nt += sg.step) {
          const int n0 = nt * SC_BN + static_cast<int>(rank) * 128;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * SP_STAGE_BYTES);
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
            uint8_t* sa = smem + stage * SP_STAGE_BYTES;
            tma_load_2d_pair(sa, &tmA, fb, kb * SC_BK, m0);
            tma_load_2d_pair(sa + SP_STAGE_BYTES / 2, &tmB, fb, kb * SC_BK, n0);
            if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        more = next_seg(f, sg);
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, SC_BN, false, false);
      const bool elected = elect_one();
      const uint64_t adesc0 = umma_smem_desc(smem_u32(smem), 16, 1024);
      const uint64_t bdesc0 = umma_smem_desc(smem_u32(smem) + SP_STAGE_BYTES / 2, 16, 1024);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (long long t = 0; t < my_tiles; ++t) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t soff = static_cast<uint32_t>(stage) * (SP_STAGE_BYTES >> 4);
          if (elected) {
#pragma unroll
            for (int k = 0; k < SC_BK / 16; ++k)
              umma_bf16_pair(tmem_base + acc * SC_BN, adesc0 + soff + k * 2, bdesc0 + soff + k * 2, idesc,
                             (kb | k) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma_commit_pair(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    long long f;
    for (bool more = first_seg(f); more;) {
      const Seg sg = get_seg(f);
      const int row = sg.mt * 256 + static_cast<int>(rank) * 128 + quad * 32 + lane;
      const bool row_ok = row < p.B;
      TopkState st;
      if (MODE == SC_TOPK) topk_state_init(st, s_list, threadIdx.x - 64, p, row, row_ok);
      for (int nt = sg.lo; nt < sg.hi; nt += sg.step) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const long long n0 = static_cast<long long>(nt) * SC_BN + half * (SC_BN / 2);
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * SC_BN + half * (SC_BN / 2);
        uint32_t rbuf[2][32];
        tmem_ld32(tbase, rbuf[0]);
        if (MODE == SC_TOPK) topk_sync_threshold(st);
#pragma unroll
        for (int c = 0; c < SC_BN / 2 / 32; ++c) {
          tmem_ld_wait();
          if (c + 1 < SC_BN / 2 / 32) {
            tmem_ld32(tbase + (c + 1) * 32, rbuf[(c + 1) & 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
          }
          const long long col0 = n0 + c * 32;
          if (col0 >= p.N) continue;
          if (MODE == SC_DENSE) dense_chunk(rbuf[c & 1], p, row, row_ok, col0);
          else topk_chunk(rbuf[c & 1], st, p, col0);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (MODE == SC_TOPK && row_ok) topk_state_flush(st, p, part * 2 + half, row);
      more = next_seg(f, sg);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// Merge `parts` lists of k (score, id) per user -> global top-k (score desc, id asc).  ONE WARP PER USER ROW: lane q
// holds entry q of the running sorted list; candidates are streamed 32 at a time (coalesced), a ballot picks those that
// beat the current k-th entry, and each is inserted cooperatively (position = number of better entries by ballot/popc,
// the tail shifts by one shuffle).  The former thread-per-row version walked parts x k candidates through a serial
// register insertion in 32 CTAs: 106 us for 28 parts x 4096 users — 14 % of a 125k-item shard's pass.
// inputs: list q of (part s, row b) at scores[s * in_part_stride + b * in_ld + q] (ids likewise, label scores at
// label_scores[s * B + b] for the dense layout, at [s * in_part_stride + b * in_ld] for packed rows); outputs with row
// strides out_ld / out_label_ld.  Dense [parts][B][k]: in_ld = k, in_part_stride = B * k; a packed (B, 2k+1) row =
// k scores | k ids | label.
// Fan-out of the merged rows into the symmetric "gathered" buffers of all ranks (rf_cosine_topk_bcast): base[r] is
// rank r's (world, B, 2k+1) buffer mapped into this process (NVLink peer memory), slot = this rank's (B, 2k+1) block in it.
struct MergeFanout {
  float* base[RF_MAX_PEERS];
  int n;
  long long slot;      // element offset of this rank's block
};

__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids,
                  const float* __restrict__ label_scores, int parts, int B, int k,
                  float* __restrict__ out_scores, int32_t* __restrict__ out_ids,
                  float* __restrict__ out_label, int in_ld, int in_part_stride, int out_ld, int out_label_ld,
                  const MergeFanout fan = MergeFanout{}) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;                                   // warp-uniform
  float ts = -INFINITY;
  int ti = 0x7fffffff;
  const int total = parts * k;
  for (int base = 0; base < total; base += 32) {
    const int idx = base + lane;
    const bool in = idx < total;
    const int sp = in ? idx / k : 0, q0 = in ? idx % k : 0;
    const size_t off = static_cast<size_t>(sp) * in_part_stride + static_cast<size_t>(b) * in_ld + q0;
    const float cs = in ? scores[off] : -INFINITY;
    const int ci = in ? ids[off] : 0x7fffffff;
    const float thr_s = __shfl_sync(0xffffffffu, ts, k - 1);
    const int thr_i = __shfl_sync(0xffffffffu, ti, k - 1);
    // -inf = list slot never filled, NaN = part never written for this row: neither is a candidate
    const bool cand = in && (cs > -INFINITY) && (cs > thr_s || (cs == thr_s && ci < thr_i));
    unsigned m = __ballot_sync(0xffffffffu, cand);
    while (m != 0) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float c_s = __shfl_sync(0xffffffffu, cs, src);
      const int c_i = __shfl_sync(0xffffffffu, ci, src);
      const bool better = lane < k && (ts > c_s || (ts == c_s && ti < c_i));
      const int pos = __popc(__ballot_sync(0xffffffffu, better));
      const float up_s = __shfl_up_sync(0xffffffffu, ts, 1);
      const int up_i = __shfl_up_sync(0xffffffffu, ti, 1);
      if (pos < k) {
        if (lane == pos) { ts = c_s; ti = c_i; }
        else if (lane > pos && lane < k) { ts = up_s; ti = up_i; }
      }
    }
  }
  float lab = -INFINITY;
  if (label_scores != nullptr)
    for (int sp = lane; sp < parts; sp += 32)
      lab = fmaxf(lab, in_ld == k ? label_scores[static_cast<size_t>(sp) * B + b]
                                  : label_scores[static_cast<size_t>(sp) * in_part_stride + static_cast<size_t>(b) * in_ld]);
  lab = warp_max(lab);
  if (fan.n > 0) {
    // one packed row (k scores | k ids | label score) per user, stored straight into every rank's gathered buffer:
    // lane q < k carries score q, lane k + q id q, lane 2k the label score — one coalesced (2k+1)-word store per peer
    const int ld = 2 * k + 1;
    const float id_bits = __int_as_float(__shfl_sync(0xffffffffu, ti, lane >= k ? lane - k : 0));
    const float word = lane < k ? ts : (lane < 2 * k ? id_bits : lab);
    if (lane < ld)
      for (int r = 0; r < fan.n; ++r) fan.base[r][fan.slot + static_cast<long long>(b) * ld + lane] = word;
    return;
  }
  if (lane < k) {
    out_scores[static_cast<size_t>(b) * out_ld + lane] = ts;
    out_ids[static_cast<size_t>(b) * out_ld + lane] = ti;
  }
  if (out_label != nullptr && lane == 0) out_label[static_cast<size_t>(b) * out_label_ld] = lab;
}

// y[n,:] = x[n,:] / max(|x[n,:]|, 1e-8) as bf16; one warp per row (E = 768).
template <bool IN_BF16>
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const void* __restrict__ x_, __nv_bfloat16* __restrict__ y, float* __restrict__ norms,
                      long long N) {
  constexpr int E = 768;
  const int lane = threadIdx.x & 31;
  for (long long n = blockIdx.x * 8ll + (threadIdx.x >> 5); n < N; n += gridDim.x * 8ll) {
    float v[24];
    if (IN_BF16) {
      const uint2* xr = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x_) + n * E);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const uint2 raw = xr[k * 32 + lane];
        const float2 a = unpack_bf16(raw.x), b = unpack_bf16(raw.y);
        v[k * 4] = a.x; v[k * 4 + 1] = a.y; v[k * 4 + 2] = b.x; v[k * 4 + 3] = b.y;
      }
    } else {
      const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x_) + n * E);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float4 a = xr[k * 32 + lane];
        v[k * 4] = a.x; v[k * 4 + 1] = a.y; v[k * 4 + 2] = a.z; v[k * 4 + 3] = a.w;
      }
    }
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) q += v[i] * v[i];
    const float nrm = sqrtf(warp_sum(q));
    const float inv = 1.0f / fmaxf(nrm, 1e-8f);
    if (norms && lane == 0) norms[n] = nrm;
    uint2* yr = reinterpret_cast<uint2*>(y + n * E);
#pragma unroll
    for (int k = 0; k < 6; ++k)
      yr[k * 32 + lane] = make_uint2(pack_bf16(v[k * 4] * inv, v[k * 4 + 1] * inv),
                                     pack_bf16(v[k * 4 + 2] * inv, v[k * 4 + 3] * inv));
  }
}

// ---- cross entropy over dense logits + gradient w.r.t. the (un-normalised) pooled vector ----
// ce_row_kernel: grid (B): loss_b, and dlogits (in place) = (softmax - onehot) / B
__global__ void __launch_bounds__(256)
ce_row_kernel(float* __restrict__ logits, const int64_t* __restrict__ labels, int B, long long N,
              float* __restrict__ loss) {
  const int b = blockIdx.x;
  float* row = logits + static_cast<size_t>(b) * N;
  __shared__ float red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY;
  for (long long n = tid; n < N; n += 256) m = fmaxf(m, row[n]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float l = 0.f;
  for (long long n = tid; n < N; n += 256) l += expf(row[n] - m);
  l = warp_sum(l);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) l += red[w];
  const long long lab = labels[b];
  const bool lab_ok = lab >= 0 && lab < N;   // out of range: the loss becomes NaN (torch device-asserts here)
  const float lse = m + logf(l);
  if (tid == 0) atomicAdd(loss, lab_ok ? (lse - row[lab]) / B : NAN);
  __syncthreads();
  const float invB = 1.0f / B;
  for (long long n = tid; n < N; n += 256) {
    const float pr = expf(row[n] - lse);
    row[n] = (pr - (n == lab ? 1.f : 0.f)) * invB;
  }
}

// Masked-LM cross entropy over dense logits [M, ld] fp32 (vocab V <= ld, padded columns ignored):
//   loss += (lse_row - logit[row, label]) / count   for rows with label >= 0 (ignore_index -100 otherwise)
//   dlogits[row, :] = (softmax - onehot) / count as bf16 (zero for ignored rows and for padded columns)
// count = number of rows with a valid label (device scalar).  grid (M), 256 threads, 3 passes over the row.
__global__ void __launch_bounds__(256)
mlm_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int V, long long ld,
              const float* __restrict__ count, float* __restrict__ loss, __nv_bfloat16* __restrict__ dlogits) {
  const long long row_off = static_cast<long long>(blockIdx.x) * ld;
  const float* row = logits + row_off;
  __nv_bfloat16* drow = dlogits + row_off;
  const long long lab = labels[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (lab < 0 || lab >= V) {   // ignore_index (negative) row; a label beyond the vocabulary poisons the loss
    for (long long n = tid * 8; n < ld; n += 256 * 8) *reinterpret_cast<uint4*>(drow + n) = make_uint4(0, 0, 0, 0);
    if (lab >= V && tid == 0) atomicAdd(loss, NAN);
    return;
  }
  __shared__ float red[8];
  float m = -INFINITY;
  for (int n = tid; n < V; n += 256) m = fmaxf(m, row[n]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float l = 0.f;
  for (int n = tid; n < V; n += 256) l += __expf(row[n] - m);
  l = warp_sum(l);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) l += red[w];
  const float lse = m + logf(l);
  const float inv = 1.0f / fmaxf(*count, 1.0f);
  if (tid == 0) atomicAdd(loss, (lse - row[lab]) * inv);
  for (long long n0 = tid * 8; n0 < ld; n0 += 256 * 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const long long n = n0 + e;
      v[e] = n < V ? (__expf(row[n] - lse) - (n == lab ? 1.f : 0.f)) * inv : 0.f;
    }
    *reinterpret_cast<uint4*>(drow + n0) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

// dxn[b,:] += inv_temp * sum_{n in chunk} dlogit[b,n] * yn[n,:]   grid (chunks, B), 256 threads
__global__ void __launch_bounds__(256)
ce_dxn_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ yn, long long N, int chunk,
              float inv_temp, float* __restrict__ dxn) {
  const int b = blockIdx.y;
  const long long n0 = static_cast<long long>(blockIdx.x) * chunk;
  const long long n1 = (n0 + chunk < N) ? n0 + chunk : N;
  const int tid = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* dl = dlogits + static_cast<size_t>(b) * N;
  for (long long n = n0; n < n1; ++n) {
    const float g = dl[n];
    const uint32_t* yr = reinterpret_cast<const uint32_t*>(yn + n * 768);
    const float2 v0 = unpack_bf16(yr[tid]);
    a0 += g * v0.x; a1 += g * v0.y;
    if (tid < 128) {
      const float2 v1 = unpack_bf16(yr[256 + tid]);
      a2 += g * v1.x; a3 += g * v1.y;
    }
  }
  float* o = dxn + static_cast<size_t>(b) * 768;
  atomicAdd(o + tid * 2, a0 * inv_temp);
  atomicAdd(o + tid * 2 + 1, a1 * inv_temp);
  if (tid < 128) {
    atomicAdd(o + 512 + tid * 2, a2 * inv_temp);
    atomicAdd(o + 512 + tid * 2 + 1, a3 * inv_temp);
  }
}

// ---- candidate scoring (ref: recformer/models.py:539-545 with `candidates`, :593-597 sampled softmax) ----
// logits[b,c] = (x_b / |x_b|) . yn[cand[b,c]] / temp: one warp per (b, c) gathers a bf16 table row (E = 768).
__global__ void __launch_bounds__(256)
cand_logits_kernel(const float* __restrict__ pooled, const __nv_bfloat16* __restrict__ yn,
                   const int64_t* __restrict__ cand, int C, long long N, float inv_temp, float* __restrict__ logits) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* xr = pooled + static_cast<size_t>(b) * 768;
  float xv[24];
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(xr + (k * 32 + lane) * 8);
    const float4 c = *reinterpret_cast<const float4*>(xr + (k * 32 + lane) * 8 + 4);
    xv[k * 8 + 0] = a.x; xv[k * 8 + 1] = a.y; xv[k * 8 + 2] = a.z; xv[k * 8 + 3] = a.w;
    xv[k * 8 + 4] = c.x; xv[k * 8 + 5] = c.y; xv[k * 8 + 6] = c.z; xv[k * 8 + 7] = c.w;
#pragma unroll
    for (int e = 0; e < 8; ++e) q += xv[k * 8 + e] * xv[k * 8 + e];
  }
  const float scale = inv_temp / fmaxf(sqrtf(warp_sum(q)), 1e-8f);
  for (int c = blockIdx.x * 8 + warp; c < C; c += gridDim.x * 8) {
    long long id = cand[static_cast<size_t>(b) * C + c];
    const bool id_ok = id >= 0 && id < N;     // out-of-range candidate: NaN logit (never a silent clamp)
    id = id < 0 ? 0 : (id >= N ? N - 1 : id);
    const uint4* yr = reinterpret_cast<const uint4*>(yn + id * 768);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint4 raw = yr[k * 32 + lane];
      const float2 y0 = unpack_bf16(raw.x), y1 = unpack_bf16(raw.y), y2 = unpack_bf16(raw.z), y3 = unpack_bf16(raw.w);
      acc += xv[k * 8 + 0] * y0.x + xv[k * 8 + 1] * y0.y + xv[k * 8 + 2] * y1.x + xv[k * 8 + 3] * y1.y +
             xv[k * 8 + 4] * y2.x + xv[k * 8 + 5] * y2.y + xv[k * 8 + 6] * y3.x + xv[k * 8 + 7] * y3.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) logits[static_cast<size_t>(b) * C + c] = id_ok ? acc * scale : NAN;
  }
}

// dxn[b,:] = inv_temp * sum_c dlogit[b,c] * yn[cand[b,c],:]   grid (B), 256 threads (3 columns each)
__global__ void __launch_bounds__(256)
cand_dxn_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ yn, const int64_t* __restrict__ cand,
                int C, long long N, float inv_temp, float* __restrict__ dxn) {
  const int b = blockIdx.x, tid = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int c = 0; c < C; ++c) {
    long long id = cand[static_cast<size_t>(b) * C + c];
    id = id < 0 ? 0 : (id >= N ? N - 1 : id);
    const float g = dlogits[static_cast<size_t>(b) * C + c];
    const __nv_bfloat16* yr = yn + id * 768;
    a0 += g * __bfloat162float(yr[tid]);
    a1 += g * __bfloat162float(yr[256 + tid]);
    a2 += g * __bfloat162float(yr[512 + tid]);
  }
  float* o = dxn + static_cast<size_t>(b) * 768;
  o[tid] = a0 * inv_temp; o[256 + tid] = a1 * inv_temp; o[512 + tid] = a2 * inv_temp;
}

// dx = (dxn - xh (xh . dxn)) / max(|x|, eps);  grid (B), 256 threads (3 elements each)
template <bool IN_BF16>
__global__ void __launch_bounds__(256)
ce_norm_bwd_kernel(const void* __restrict__ x_, const float* __restrict__ dxn, float* __restrict__ dx) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ float red[2][8];
  float xv[3], gv[3];
  float q = 0.f, d = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int e = k * 256 + tid;
    xv[k] = IN_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x_)[static_cast<size_t>(b) * 768 + e])
                    : reinterpret_cast<const float*>(x_)[static_cast<size_t>(b) * 768 + e];
    gv[k] = dxn[static_cast<size_t>(b) * 768 + e];
    q += xv[k] * xv[k];
    d += xv[k] * gv[k];
  }
  q = warp_sum(q); d = warp_sum(d);
  if (lane == 0) { red[0][warp] = q; red[1][warp] = d; }
  __syncthreads();
  q = 0.f; d = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { q += red[0][w]; d += red[1][w]; }
  const float nrm = fmaxf(sqrtf(q), 1e-8f);
  const float inv = 1.f / nrm;
  // xh = x*inv; xh.dxn = d*inv
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int e = k * 256 + tid;
    dx[static_cast<size_t>(b) * 768 + e] = (gv[k] - xv[k] * inv * (d * inv)) * inv;
  }
}

static bool use_pair(int B) { return B > SC_BM; }

// Fills slices / extra / main_tiles / extra_tiles; returns the number of (slice | leftover pair) parts.
static int plan_schedule(int B, long long N, ScoreParams* p) {
  const long long n_tiles = (N + SC_BN - 1) / SC_BN;
  long long s;
  int extra = 0;
  long long main_tiles = n_tiles, extra_tiles = 0, extra_steps = 0;
  if (use_pair(B)) {
    const int pairs = sm_count() / 2, m_tiles = (B + 255) / 256;
    s = pairs / m_tiles;
    if (s < 1) s = 1;
    if (s > n_tiles) s = n_tiles;
    extra = pairs - static_cast<int>(s) * m_tiles;
    if (extra > 0 && n_tiles >= 4ll * pairs) {
      // a leftover pair restarts its top-k list at every segment and so spends more time inserting than a
      // lock-step pair: it also streams its table range from HBM rather than L2.  Measured (4096 x 1M): giving it ~75 % of a
      // lock-step pair's steps minimises the pass time
      static const int pct = getenv("RF_SCORE_EXTRA_PCT") ? atoi(getenv("RF_SCORE_EXTRA_PCT")) : 75;   // tuning aid
      extra_tiles = n_tiles * extra * pct / (100ll * pairs);
      if (extra_tiles == 0) {
        extra = 0;
      } else {
        main_tiles = n_tiles - extra_tiles;
        extra_steps = (extra_tiles * m_tiles + extra - 1) / extra;
      }
    } else {
      extra = 0;
    }
  } else {
    s = sm_count() / ((B + SC_BM - 1) / SC_BM);
    if (s < 1) s = 1;
    if (s > n_tiles) s = n_tiles;
  }
  if (p) {
    p->slices = static_cast<int>(s); p->extra = extra;
    p->main_tiles = static_cast<int>(main_tiles); p->extra_tiles = static_cast<int>(extra_tiles);
    p->extra_steps = static_cast<int>(extra_steps);
  }
  return static_cast<int>(s) + extra;
}

static int launch_cosine(int mode, const void* xn, const void* yn, ScoreParams& p, cudaStream_t stream) {
  static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
  if (first_use_on_device(&attr_seen)) {
    RF_CUDA(cudaFuncSetAttribute(cosine_mma_kernel<SC_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
    RF_CUDA(cudaFuncSetAttribute(cosine_mma_kernel<SC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM));
    RF_CUDA(cudaFuncSetAttribute(cosine_pair_kernel<SC_TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
    RF_CUDA(cudaFuncSetAttribute(cosine_pair_kernel<SC_DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
  }
  if (use_pair(p.B)) {
    const CUtensorMap* tmA = get_tmap_2d(xn, p.B, p.K, p.K, 128);
    const CUtensorMap* tmB = get_tmap_2d(yn, static_cast<uint64_t>(p.N), p.K, p.K, 128);
    if (!tmA || !tmB) return RF_ERR_CUDA;
    const int grid = 2 * (((p.B + 255) / 256) * p.slices + p.extra);
    if (mode == SC_TOPK)
      cosine_pair_kernel<SC_TOPK><<<grid, SC_THREADS, SP_SMEM, stream>>>(*tmA, *tmB, p);
    else
      cosine_pair_kernel<SC_DENSE><<<grid, SC_THREADS, SP_SMEM, stream>>>(*tmA, *tmB, p);
    return check_launch("cosine_pair_kernel");
  }
  const CUtensorMap* tmA = get_tmap_2d(xn, p.B, p.K, p.K, SC_BM);
  const CUtensorMap* tmB = get_tmap_2d(yn, static_cast<uint64_t>(p.N), p.K, p.K, SC_BN);
  if (!tmA || !tmB) return RF_ERR_CUDA;
  const int grid = ((p.B + SC_BM - 1) / SC_BM) * p.slices;
  if (mode == SC_TOPK)
    cosine_mma_kernel<SC_TOPK><<<grid, SC_THREADS, SC_SMEM, stream>>>(*tmA, *tmB, p);
  else
    cosine_mma_kernel<SC_DENSE><<<grid, SC_THREADS, SC_SMEM, stream>>>(*tmA, *tmB, p);
  return check_launch("cosine_mma_kernel");
}

}  // namespace rf

using namespace rf;

extern "C" int rf_normalize_rows(const void* x, int x_is_bf16, void* y, float* norms, long long N, int E,
                                 rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(x && y && N > 0, "rf_normalize_rows: bad argument");
  RF_REQUIRE(E == 768, "rf_normalize_rows: hidden size %d unsupported (768)", E);
  long long grid = (N + 7) / 8;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  if (x_is_bf16)
    normalize_rows_kernel<true><<<static_cast<int>(grid), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), norms, N);
  else
    normalize_rows_kernel<false><<<static_cast<int>(grid), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), norms, N);
  return check_launch("rf_normalize_rows");
}

extern "C" int rf_cosine_logits(const void* xn, const void* yn, float* logits, int B, long long N, int E, float temp,
                                rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(xn && yn && logits && B > 0 && N > 0, "rf_cosine_logits: bad argument");
  RF_REQUIRE(E % 8 == 0, "rf_cosine_logits: E must be a multiple of 8");
  ScoreParams p{};
  p.B = B; p.N = N; p.K = E; p.inv_temp = 1.0f / temp; p.k = 0; p.id_base = 0;
  plan_schedule(B, N, &p);
  p.logits = logits;
  return launch_cosine(SC_DENSE, xn, yn, p, stream);
}

extern "C" long long rf_cosine_topk_ws_bytes(int B, long long N, int k) {
  const long long parts = 2ll * plan_schedule(B, N, nullptr);
  return parts * B * (static_cast<long long>(k) * 8 + 4) + (static_cast<long long>(B) + 1) * 4 + 256;
}

static int cosine_topk_impl(const void* xn, const void* yn, int B, long long N, int E, float temp, int k, int id_base,
                            const int64_t* labels, float* topk_scores, int32_t* topk_ids, float* label_score, int out_ld,
                            int label_ld, void* ws, cudaStream_t stream, const MergeFanout* fan = nullptr) {
  RF_REQUIRE(xn && yn && topk_scores && (topk_ids || fan) && ws && B > 0 && N > 0, "rf_cosine_topk: bad argument");
  RF_REQUIRE(static_cast<long long>(B) * (2 * k + 1) < 2147483647ll, "rf_cosine_topk: B * k too large");
  RF_REQUIRE(k >= 1 && k <= SC_MAXK, "rf_cosine_topk: k=%d out of range [1,%d]", k, SC_MAXK);
  RF_REQUIRE(E % 8 == 0, "rf_cosine_topk: E must be a multiple of 8");
  RF_REQUIRE(N + id_base < 2147483647ll, "rf_cosine_topk: item ids must fit int32");
  ScoreParams p{};
  p.B = B; p.N = N; p.K = E; p.inv_temp = 1.0f / temp; p.k = k; p.id_base = id_base;
  const int parts = 2 * plan_schedule(B, N, &p);   // every (slice | leftover pair, 128-column half) keeps its own list
  p.labels = labels;
  const size_t cnt = static_cast<size_t>(parts) * B * k;
  p.ws_scores = reinterpret_cast<float*>(ws);
  p.ws_ids = reinterpret_cast<int32_t*>(p.ws_scores + cnt);
  p.ws_label = reinterpret_cast<float*>(p.ws_ids + cnt);
  // a leftover pair only writes the rows of the user tiles it visited: every other (part, row) slot keeps
  // this NaN pattern, which the merge skips
  p.ws_thr = reinterpret_cast<unsigned int*>(p.ws_label + static_cast<size_t>(parts) * B);
  RF_CUDA(cudaMemsetAsync(p.ws_thr, 0, (static_cast<size_t>(B) + 1) * 4, stream));     // 0 = no threshold published yet
  // (Seeding the shared thresholds from a pre-pass over the first 4096 items was measured: the pre-pass costs 88 us and
  // the main sweep gains nothing, because with 32 rows per warp a chunk is only skipped once the threshold sits near the
  // 1e-4 quantile.)
  int rc;
  RF_CUDA(cudaMemsetAsync(ws, 0xFF, cnt * 8 + static_cast<size_t>(parts) * B * 4, stream));
  rc = launch_cosine(SC_TOPK, xn, yn, p, stream);
  if (rc) return rc;
  topk_merge_kernel<<<(B + 7) / 8, 256, 0, stream>>>(p.ws_scores, p.ws_ids, p.ws_label, parts, B, k, topk_scores,
                                                        topk_ids, label_score, k, B * k, out_ld, label_ld,
                                                        fan ? *fan : MergeFanout{});
  return check_launch("rf_cosine_topk/merge");
}

extern "C" int rf_cosine_topk(const void* xn, const void* yn, int B, long long N, int E, float temp, int k,
                              int id_base, const int64_t* labels, float* topk_scores, int32_t* topk_ids,
                              float* label_score, void* ws, rf_stream_t stream_) {
  RF_REQUIRE(topk_scores && topk_ids, "rf_cosine_topk: bad argument");
  return cosine_topk_impl(xn, yn, B, N, E, temp, k, id_base, labels, topk_scores, topk_ids, label_score, k, 1, ws,
                          reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int rf_cosine_topk_packed(const void* xn, const void* yn, int B, long long N, int E, float temp, int k,
                                     int id_base, const int64_t* labels, float* packed, void* ws, rf_stream_t stream_) {
  RF_REQUIRE(packed, "rf_cosine_topk_packed: bad argument");
  const int ld = 2 * k + 1;
  return cosine_topk_impl(xn, yn, B, N, E, temp, k, id_base, labels, packed, reinterpret_cast<int32_t*>(packed) + k,
                          packed + 2 * k, ld, ld, ws, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int rf_cosine_topk_bcast(const void* xn, const void* yn, int B, long long N, int E, float temp, int k,
                                    int id_base, const int64_t* labels, const unsigned long long* peer_gathered, int world,
                                    int rank, void* ws, rf_stream_t stream_) {
  RF_REQUIRE(peer_gathered && world >= 1 && world <= RF_MAX_PEERS && rank >= 0 && rank < world,
             "rf_cosine_topk_bcast: bad peer list (world=%d, rank=%d, at most %d peers)", world, rank, RF_MAX_PEERS);
  RF_REQUIRE(2 * k + 1 <= 32, "rf_cosine_topk_bcast: k=%d too large for one packed row per warp", k);
  MergeFanout fan{};
  fan.n = world;
  fan.slot = static_cast<long long>(rank) * B * (2 * k + 1);
  for (int r = 0; r < world; ++r) {
    RF_REQUIRE(peer_gathered[r] != 0 && (peer_gathered[r] & 3) == 0, "rf_cosine_topk_bcast: null / misaligned peer buffer %d", r);
    fan.base[r] = reinterpret_cast<float*>(static_cast<uintptr_t>(peer_gathered[r]));
  }
  // the merged rows never land in a local (B, 2k+1) buffer: the merge kernel's only stores are the peer stores
  return cosine_topk_impl(xn, yn, B, N, E, temp, k, id_base, labels, fan.base[rank], nullptr, nullptr, 2 * k + 1, 2 * k + 1,
                          ws, reinterpret_cast<cudaStream_t>(stream_), &fan);
}

extern "C" int rf_topk_merge_packed(const float* packed, int parts, int B, int k, float* out_scores, int32_t* out_ids,
                                    float* out_label_score, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(packed && out_scores && out_ids && parts > 0 && B > 0, "rf_topk_merge_packed: bad argument");
  RF_REQUIRE(k >= 1 && k <= SC_MAXK, "rf_topk_merge_packed: k=%d out of range [1,%d]", k, SC_MAXK);
  const int ld = 2 * k + 1;
  topk_merge_kernel<<<(B + 7) / 8, 256, 0, stream>>>(packed, reinterpret_cast<const int32_t*>(packed) + k, packed + 2 * k,
                                                        parts, B, k, out_scores, out_ids, out_label_score, ld, B * ld, k, 1);
  return check_launch("rf_topk_merge_packed");
}

extern "C" int rf_topk_merge(const float* scores, const int32_t* ids, const float* label_scores, int parts, int B,
                             int k, float* out_scores, int32_t* out_ids, float* out_label_score,
                             rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(scores && ids && out_scores && out_ids && parts > 0 && B > 0, "rf_topk_merge: bad argument");
  RF_REQUIRE(k >= 1 && k <= SC_MAXK, "rf_topk_merge: k=%d out of range [1,%d]", k, SC_MAXK);
  topk_merge_kernel<<<(B + 7) / 8, 256, 0, stream>>>(scores, ids, label_scores, parts, B, k, out_scores, out_ids,
                                                        out_label_score, k, B * k, k, 1);
  return check_launch("rf_topk_merge");
}

extern "C" long long rf_cosine_ce_ws_bytes(int B, long long N, int E) {
  return static_cast<long long>(B) * N * 4 + static_cast<long long>(B) * E * (2 + 4) + 256;
}

// ws layout: [B*E bf16 xn][B*N fp32 logits][B*E fp32 dxn]
extern "C" int rf_cosine_ce(const void* pooled, int pooled_is_bf16, const void* yn, const int64_t* labels, int B,
                            long long N, int E, float temp, float* loss, float* dpooled, float* ws_,
                            rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(pooled && yn && labels && loss && ws_ && B > 0 && N > 0, "rf_cosine_ce: bad argument");
  RF_REQUIRE(E == 768, "rf_cosine_ce: hidden size %d unsupported (768)", E);
  uint8_t* ws = reinterpret_cast<uint8_t*>(ws_);
  __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(ws);
  size_t off = (static_cast<size_t>(B) * E * 2 + 255) & ~size_t(255);
  float* logits = reinterpret_cast<float*>(ws + off);
  off += (static_cast<size_t>(B) * N * 4 + 255) & ~size_t(255);
  float* dxn = reinterpret_cast<float*>(ws + off);
  int rc = rf_normalize_rows(pooled, pooled_is_bf16, xn, nullptr, B, E, stream_);
  if (rc) return rc;
  rc = rf_cosine_logits(xn, yn, logits, B, N, E, temp, stream_);
  if (rc) return rc;
  RF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  ce_row_kernel<<<B, 256, 0, stream>>>(logits, labels, B, N, loss);
  rc = check_launch("rf_cosine_ce/row");
  if (rc || dpooled == nullptr) return rc;
  RF_CUDA(cudaMemsetAsync(dxn, 0, static_cast<size_t>(B) * E * 4, stream));
  const int chunk = 128;
  ce_dxn_kernel<<<dim3(static_cast<unsigned>((N + chunk - 1) / chunk), B), 256, 0, stream>>>(
      logits, reinterpret_cast<const __nv_bfloat16*>(yn), N, chunk, 1.0f / temp, dxn);
  rc = check_launch("rf_cosine_ce/dxn");
  if (rc) return rc;
  if (pooled_is_bf16)
    ce_norm_bwd_kernel<true><<<B, 256, 0, stream>>>(pooled, dxn, dpooled);
  else
    ce_norm_bwd_kernel<false><<<B, 256, 0, stream>>>(pooled, dxn, dpooled);
  return check_launch("rf_cosine_ce/norm_bwd");
}

extern "C" int rf_mlm_ce(const float* logits, const int64_t* labels, int M, int V, long long ld, const float* count,
                         float* loss, void* dlogits_bf16, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(logits && labels && count && loss && dlogits_bf16 && M > 0 && V > 0, "rf_mlm_ce: bad argument");
  RF_REQUIRE(ld >= V && ld % 8 == 0, "rf_mlm_ce: ld=%lld must be >= V and a multiple of 8", ld);
  RF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  mlm_ce_kernel<<<M, 256, 0, stream>>>(logits, labels, V, ld, count, loss, reinterpret_cast<__nv_bfloat16*>(dlogits_bf16));
  return check_launch("rf_mlm_ce");
}

// logits[b,c] = cos(pooled_b, table[cand[b,c]]) / temp over the pre-normalised bf16 table.
extern "C" int rf_cosine_candidates(const float* pooled, const void* yn, const int64_t* cand, int B, int C, long long N,
                                    int E, float temp, float* logits, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(pooled && yn && cand && logits && B > 0 && C > 0 && N > 0, "rf_cosine_candidates: bad argument");
  RF_REQUIRE(E == 768, "rf_cosine_candidates: hidden size %d unsupported (768)", E);
  int gx = (C + 7) / 8;
  if (gx > 64) gx = 64;
  cand_logits_kernel<<<dim3(gx, B), 256, 0, stream>>>(pooled, reinterpret_cast<const __nv_bfloat16*>(yn), cand, C, N,
                                                       1.0f / temp, logits);
  return check_launch("rf_cosine_candidates");
}

// Sampled-softmax CE (ref: recformer/models.py:593-597): loss = mean_b CE(logits[b,:], target 0) with the label in
// column 0 of `cand`; dpooled = gradient w.r.t. the un-normalised pooled vector.  ws: B*C*4 + B*E*4 + 8*B + 256 bytes.
extern "C" int rf_cosine_candidates_ce(const float* pooled, const void* yn, const int64_t* cand, int B, int C, long long N,
                                       int E, float temp, float* loss, float* dpooled, void* ws_, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(pooled && yn && cand && loss && ws_ && B > 0 && C > 0, "rf_cosine_candidates_ce: bad argument");
  uint8_t* ws = reinterpret_cast<uint8_t*>(ws_);
  float* logits = reinterpret_cast<float*>(ws);
  size_t off = (static_cast<size_t>(B) * C * 4 + 255) & ~size_t(255);
  float* dxn = reinterpret_cast<float*>(ws + off);
  off += (static_cast<size_t>(B) * E * 4 + 255) & ~size_t(255);
  int64_t* zeros = reinterpret_cast<int64_t*>(ws + off);
  int rc = rf_cosine_candidates(pooled, yn, cand, B, C, N, E, temp, logits, stream_);
  if (rc) return rc;
  RF_CUDA(cudaMemsetAsync(zeros, 0, static_cast<size_t>(B) * 8, stream));
  RF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  ce_row_kernel<<<B, 256, 0, stream>>>(logits, zeros, B, C, loss);
  rc = check_launch("rf_cosine_candidates_ce/row");
  if (rc || dpooled == nullptr) return rc;
  cand_dxn_kernel<<<B, 256, 0, stream>>>(logits, reinterpret_cast<const __nv_bfloat16*>(yn), cand, C, N, 1.0f / temp, dxn);
  rc = check_launch("rf_cosine_candidates_ce/dxn");
  if (rc) return rc;
  ce_norm_bwd_kernel<false><<<B, 256, 0, stream>>>(pooled, dxn, dpooled);
  return check_launch("rf_cosine_candidates_ce/norm_bwd");
}
