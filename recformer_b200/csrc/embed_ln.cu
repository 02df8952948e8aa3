// Memory-bound row kernels: input preparation, fused embedding-sum + LayerNorm (+dropout),
// LayerNorm forward/backward, fp32->bf16 cast and fused AdamW.  One warp per token row, 128-bit
// coalesced accesses, warp-shuffle reductions, fp32 statistics.
#include <cuda_bf16.h>
#include <math.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

RF_DEFINE_NONCE_LOADER(embed_ln)

namespace rf {

constexpr int ROW_THREADS = 256;  // 8 warps = 8 rows per CTA pass

// row activity (rf_set_row_activity): one flag per 256 token rows; a row kernel skips rows of padding-only tiles
__device__ __forceinline__ bool row_on(const uint8_t* active, int t) { return active == nullptr || active[t >> 8] != 0; }
static const uint8_t* active_flags_for(long long rows) {
  const RowActivity& ra = row_activity();
  return (ra.flags != nullptr && ra.rows == rows) ? ra.flags : nullptr;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// keep-mask for 4 consecutive elements (group index = element index / 4)
__device__ __forceinline__ uint32_t dropout_keep4(uint64_t seed, uint64_t grp4, uint32_t thresh16) {
  const uint4 r = philox4x32(seed, grp4);
  return ((r.x & 0xFFFFu) >= thresh16 ? 1u : 0u) | ((r.x >> 16) >= thresh16 ? 2u : 0u) |
         ((r.y & 0xFFFFu) >= thresh16 ? 4u : 0u) | ((r.y >> 16) >= thresh16 ? 8u : 0u);
}

// ----------------------------------------------------------------------------------------------
// rf_prepare_inputs: one warp per sequence, ballot/popc scan over 32-token groups
// ----------------------------------------------------------------------------------------------
__global__ void prepare_inputs_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ am,
                                      const int64_t* __restrict__ gm, int B, int L, int Lp, int pad,
                                      int32_t* __restrict__ pos_ids, uint8_t* __restrict__ mask012, int* err_flag) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int running = 0;
  bool bad_global = false;
  for (int l0 = 0; l0 < Lp; l0 += 32) {
    const int l = l0 + lane;
    const bool in = l < L;
    const bool nonpad = in && ids[static_cast<size_t>(b) * L + l] != pad;
    const uint32_t bal = __ballot_sync(0xffffffffu, nonpad);
    const int incl = running + __popc(bal & (0xffffffffu >> (31 - lane)));
    if (l < Lp) {
      pos_ids[static_cast<size_t>(b) * Lp + l] = nonpad ? incl + pad : pad;
      int m = 0;
      if (in) {
        const int a = am ? static_cast<int>(am[static_cast<size_t>(b) * L + l]) : 1;
        const int g = gm ? static_cast<int>(gm[static_cast<size_t>(b) * L + l]) : 0;
        m = a * (g + 1);
        if (m > 1 && l != 0) bad_global = true;
        if (m < 0 || m > 2) bad_global = true;
      }
      mask012[static_cast<size_t>(b) * Lp + l] = static_cast<uint8_t>(m < 0 ? 0 : (m > 2 ? 2 : m));
    }
    running += __popc(bal);
  }
  if (__any_sync(0xffffffffu, bad_global) && lane == 0) atomicOr(err_flag, 1);
}

// rf_row_tile_flags: which 256-row tiles of the [B*L] token axis hold at least one real token, and the compact list of
// the 128-row attention query tiles inside them.  One CTA; warp w takes tiles w, w + 8, ...
__global__ void __launch_bounds__(256) row_tile_flags_kernel(const uint8_t* __restrict__ mask012, int n_tiles,
                                                            uint8_t* __restrict__ flags, int32_t* __restrict__ qtiles,
                                                            int32_t* __restrict__ n_qtiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < n_tiles; t += 8) {
    const unsigned long long v = reinterpret_cast<const unsigned long long*>(mask012 + static_cast<size_t>(t) * 256)[lane];
    const bool any = __any_sync(0xffffffffu, v != 0ull);
    if (lane == 0) flags[t] = any ? 1 : 0;
  }
  __syncthreads();
  if (warp == 0) {
    int count = 0;
    for (int t0 = 0; t0 < n_tiles; t0 += 32) {
      const int t = t0 + lane;
      const bool on = t < n_tiles && flags[t] != 0;
      const unsigned bal = __ballot_sync(0xffffffffu, on);
      if (on) {
        const int at = count + __popc(bal & ((1u << lane) - 1u));
        qtiles[2 * at] = 2 * t;
        qtiles[2 * at + 1] = 2 * t + 1;
      }
      count += __popc(bal);
    }
    if (lane == 0) *n_qtiles = 2 * count;
  }
}

// ----------------------------------------------------------------------------------------------
// fused embedding sum + LayerNorm.  NV4 = E / 128 float4 per lane.
// ----------------------------------------------------------------------------------------------
struct EmbedDev {
  const int64_t* ids; const int64_t* tt; const int64_t* ip; const int32_t* pos;
  const float* word; const float* posw; const float* type; const float* item;
  const float* gamma; const float* beta;
  int B, L, Lp, E, vocab, max_pos, type_size, max_item, pad;
  float eps, drop_scale; uint32_t drop_thresh; uint64_t drop_seed;
};

__device__ __forceinline__ void embed_token_ids(const EmbedDev& a, int t, int& id, int& pid, int& tt, int& ip,
                                                bool& bad) {
  const int b = t / a.Lp, l = t % a.Lp;
  if (l < a.L) {
    const size_t s = static_cast<size_t>(b) * a.L + l;
    id = static_cast<int>(a.ids[s]);
    tt = a.tt ? static_cast<int>(a.tt[s]) : 0;
    ip = static_cast<int>(a.ip[s]);
  } else {  // window padding (ref: recformer/models.py:238-258)
    id = a.pad; tt = 0; ip = a.pad;
  }
  pid = a.pos[t];
  bad = id < 0 || id >= a.vocab || pid < 0 || pid >= a.max_pos || tt < 0 || tt >= a.type_size || ip < 0 ||
        ip >= a.max_item;
  id = min(max(id, 0), a.vocab - 1);
  pid = min(max(pid, 0), a.max_pos - 1);
  tt = min(max(tt, 0), a.type_size - 1);
  ip = min(max(ip, 0), a.max_item - 1);
}

template <int NV4>
__device__ __forceinline__ void embed_row_stats(const EmbedDev& a, int id, int pid, int tt, int ip, int lane,
                                                float4 (&x)[NV4], float& mean, float& rstd) {
  const float4* w = reinterpret_cast<const float4*>(a.word + static_cast<size_t>(id) * a.E);
  const float4* pw = reinterpret_cast<const float4*>(a.posw + static_cast<size_t>(pid) * a.E);
  const float4* tw = reinterpret_cast<const float4*>(a.type + static_cast<size_t>(tt) * a.E);
  const float4* iw = reinterpret_cast<const float4*>(a.item + static_cast<size_t>(ip) * a.E);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV4; ++k) {
    const int v = k * 32 + lane;
    const float4 a0 = __ldg(w + v), a1 = __ldg(pw + v), a2 = __ldg(tw + v), a3 = __ldg(iw + v);
    // same association order as the reference: ((word + pos) + type) + item
    x[k].x = ((a0.x + a1.x) + a2.x) + a3.x;
    x[k].y = ((a0.y + a1.y) + a2.y) + a3.y;
    x[k].z = ((a0.z + a1.z) + a2.z) + a3.z;
    x[k].w = ((a0.w + a1.w) + a2.w) + a3.w;
    s += (x[k].x + x[k].y) + (x[k].z + x[k].w);
  }
  mean = warp_sum(s) / a.E;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV4; ++k) {
    const float dx = x[k].x - mean, dy = x[k].y - mean, dz = x[k].z - mean, dw = x[k].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  rstd = rsqrtf(warp_sum(q) / a.E + a.eps);
}

// Raw (unvalidated) ids of one token: loaded one to two work items ahead of their use so that the dependent
// row gathers never wait for them.
struct TokIds { int id, pid, tt, ip; };

__device__ __forceinline__ TokIds embed_load_ids(const EmbedDev& a, int b, int l) {
  TokIds r;
  if (l < a.L) {
    const size_t s = static_cast<size_t>(b) * a.L + l;
    r.id = static_cast<int>(__ldg(a.ids + s));
    r.tt = a.tt ? static_cast<int>(__ldg(a.tt + s)) : 0;
    r.ip = static_cast<int>(__ldg(a.ip + s));
  } else {  // window padding (ref: recformer/models.py:238-258)
    r.id = a.pad; r.tt = 0; r.ip = a.pad;
  }
  r.pid = __ldg(a.pos + static_cast<size_t>(b) * a.Lp + l);
  return r;
}

__device__ __forceinline__ bool embed_clamp_ids(const EmbedDev& a, TokIds& r) {
  const bool bad = r.id < 0 || r.id >= a.vocab || r.pid < 0 || r.pid >= a.max_pos || r.tt < 0 || r.tt >= a.type_size ||
                   r.ip < 0 || r.ip >= a.max_item;
  r.id = min(max(r.id, 0), a.vocab - 1);
  r.pid = min(max(r.pid, 0), a.max_pos - 1);
  r.tt = min(max(r.tt, 0), a.type_size - 1);
  r.ip = min(max(r.ip, 0), a.max_item - 1);
  return bad;
}

// Work decomposition of the backward kernel: a work item is (sequence group, position l); the warps of a CTA take
// the sequences of the group at the SAME position, so that phase B (below) can sum their position-row gradients in
// registers.  Each CTA owns a contiguous range of work items (consecutive positions: the id loads of the next item
// hit the same cache lines, and a CTA touches few item-position rows).
// warp -> (sequence, position) of work item w.  `seqs` (the largest power of two <= min(B, warps)) sequences share a
// CTA; the remaining factor of the CTA's warps covers consecutive positions.
struct EmbedMap {
  int seqs, posn, lblocks, items;
  __host__ __device__ __forceinline__ void locate(int w, int warp, int& b, int& l) const {
    const int g = w / lblocks, lb = w - g * lblocks;
    b = g * seqs + (warp & (seqs - 1));
    l = lb * posn + warp / seqs;
  }
};

__host__ __device__ __forceinline__ EmbedMap embed_map(int B, int Lp, int warps) {
  EmbedMap m;
  m.seqs = 1;
  while (m.seqs * 2 <= B && m.seqs * 2 <= warps) m.seqs *= 2;
  m.posn = warps / m.seqs;
  m.lblocks = (Lp + m.posn - 1) / m.posn;
  m.items = ((B + m.seqs - 1) / m.seqs) * m.lblocks;
  return m;
}

// Forward: one warp per token, the 8 warps of a CTA on 8 consecutive tokens.  Measured alternatives (ncu, C2 shape,
// 30.4 us for this one): warps on the same position of 8 / 16 sequences so that the position row is shared (36 us: the
// 3 MB-strided output rows), token-type / item-position tables in shared memory (38 us: every CTA first pulls the same
// 165 KB through L2) — the kernel moves 16.7 KB of L2 traffic per token for 7.7 KB of compulsory bytes and sits at
// the L2 fabric's ~7.5 TB/s.
template <int NV4>
__global__ void __launch_bounds__(ROW_THREADS) embed_ln_fwd_kernel(const EmbedDev a, __nv_bfloat16* __restrict__ out,
                                                                   float* __restrict__ out32, int* err_flag,
                                                                   const uint8_t* __restrict__ active) {
  const int lane = threadIdx.x & 31;
  const int T = a.B * a.Lp;
  for (int t = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5); t < T; t += gridDim.x * (ROW_THREADS / 32)) {
    if (!row_on(active, t)) continue;
    int id, pid, tt, ip;
    bool bad;
    embed_token_ids(a, t, id, pid, tt, ip, bad);
    if (bad && lane == 0 && err_flag) atomicOr(err_flag, 2);
    float4 x[NV4];
    float mean, rstd;
    embed_row_stats<NV4>(a, id, pid, tt, ip, lane, x, mean, rstd);
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
    const float4* b4 = reinterpret_cast<const float4*>(a.beta);
    uint2* o2 = reinterpret_cast<uint2*>(out + static_cast<size_t>(t) * a.E);
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int v = k * 32 + lane;
      const float4 g = __ldg(g4 + v), be = __ldg(b4 + v);
      float y0 = (x[k].x - mean) * rstd * g.x + be.x, y1 = (x[k].y - mean) * rstd * g.y + be.y;
      float y2 = (x[k].z - mean) * rstd * g.z + be.z, y3 = (x[k].w - mean) * rstd * g.w + be.w;
      if (a.drop_thresh != 0) {
        const uint32_t keep = dropout_keep4(a.drop_seed, static_cast<uint64_t>(t) * (a.E / 4) + v, a.drop_thresh);
        y0 = (keep & 1u) ? y0 * a.drop_scale : 0.f; y1 = (keep & 2u) ? y1 * a.drop_scale : 0.f;
        y2 = (keep & 4u) ? y2 * a.drop_scale : 0.f; y3 = (keep & 8u) ? y3 * a.drop_scale : 0.f;
      }
      if (out) o2[v] = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
      if (out32) reinterpret_cast<float4*>(out32 + static_cast<size_t>(t) * a.E)[v] = make_float4(y0, y1, y2, y3);
    }
  }
}

// Fallback backward (small tables do not fit shared memory): one warp per token, every table updated by global
// red.adds — 4 x 768 atomics per token, of which the token-type / item-position ones pile onto a handful of rows.
template <int NV4>
__global__ void __launch_bounds__(ROW_THREADS)
embed_ln_bwd_atomic_kernel(const EmbedDev a, const __nv_bfloat16* __restrict__ dout, float* d_word, float* d_pos,
                    float* d_type, float* d_item, float* d_gamma, float* d_beta, const uint8_t* __restrict__ active) {
  const int lane = threadIdx.x & 31;
  const int T = a.B * a.Lp;
  float4 dg[NV4], db[NV4];
#pragma unroll
  for (int k = 0; k < NV4; ++k) dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5); t < T; t += gridDim.x * (ROW_THREADS / 32)) {
    if (!row_on(active, t)) continue;
    int id, pid, tt, ip;
    bool bad;
    embed_token_ids(a, t, id, pid, tt, ip, bad);
    float4 x[NV4];
    float mean, rstd;
    embed_row_stats<NV4>(a, id, pid, tt, ip, lane, x, mean, rstd);
    const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
    const uint2* d2 = reinterpret_cast<const uint2*>(dout + static_cast<size_t>(t) * a.E);
    float4 gy[NV4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int v = k * 32 + lane;
      const uint2 raw = d2[v];
      float2 lo = unpack_bf16(raw.x), hi = unpack_bf16(raw.y);
      float4 dy = make_float4(lo.x, lo.y, hi.x, hi.y);
      if (a.drop_thresh != 0) {
        const uint32_t keep = dropout_keep4(a.drop_seed, static_cast<uint64_t>(t) * (a.E / 4) + v, a.drop_thresh);
        dy.x = (keep & 1u) ? dy.x * a.drop_scale : 0.f; dy.y = (keep & 2u) ? dy.y * a.drop_scale : 0.f;
        dy.z = (keep & 4u) ? dy.z * a.drop_scale : 0.f; dy.w = (keep & 8u) ? dy.w * a.drop_scale : 0.f;
      }
      const float4 xh = make_float4((x[k].x - mean) * rstd, (x[k].y - mean) * rstd, (x[k].z - mean) * rstd,
                                    (x[k].w - mean) * rstd);
      x[k] = xh;
      dg[k].x += dy.x * xh.x; dg[k].y += dy.y * xh.y; dg[k].z += dy.z * xh.z; dg[k].w += dy.w * xh.w;
      db[k].x += dy.x; db[k].y += dy.y; db[k].z += dy.z; db[k].w += dy.w;
      const float4 g = __ldg(g4 + v);
      gy[k] = make_float4(dy.x * g.x, dy.y * g.y, dy.z * g.z, dy.w * g.w);
      s1 += (gy[k].x + gy[k].y) + (gy[k].z + gy[k].w);
      s2 += (gy[k].x * xh.x + gy[k].y * xh.y) + (gy[k].z * xh.z + gy[k].w * xh.w);
    }
    s1 = warp_sum(s1) / a.E;
    s2 = warp_sum(s2) / a.E;
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      const int v = k * 32 + lane;
      const float dx0 = rstd * (gy[k].x - s1 - x[k].x * s2), dx1 = rstd * (gy[k].y - s1 - x[k].y * s2);
      const float dx2 = rstd * (gy[k].z - s1 - x[k].z * s2), dx3 = rstd * (gy[k].w - s1 - x[k].w * s2);
      // nn.Embedding(padding_idx=pad) never accumulates a gradient into the pad row
      // (ref: recformer/models.py:89,104-106)
      if (d_word && id != a.pad) red_add_v4(d_word + static_cast<size_t>(id) * a.E + v * 4, dx0, dx1, dx2, dx3);
      if (d_pos && pid != a.pad) red_add_v4(d_pos + static_cast<size_t>(pid) * a.E + v * 4, dx0, dx1, dx2, dx3);
      if (d_type) red_add_v4(d_type + static_cast<size_t>(tt) * a.E + v * 4, dx0, dx1, dx2, dx3);
      if (d_item) red_add_v4(d_item + static_cast<size_t>(ip) * a.E + v * 4, dx0, dx1, dx2, dx3);
    }
  }
#pragma unroll
  for (int k = 0; k < NV4; ++k) {
    const int v = k * 32 + lane;
    if (d_gamma) red_add_v4(d_gamma + v * 4, dg[k].x, dg[k].y, dg[k].z, dg[k].w);
    if (d_beta) red_add_v4(d_beta + v * 4, db[k].x, db[k].y, db[k].z, db[k].w);
  }
}

// Backward with CTA-level aggregation of the small tables.  Same work decomposition as the forward (the warps of a
// CTA take up to 8 sequences at the same position).  Per work item:
//   phase A (warp = token): recompute the row statistics, dx = LayerNorm backward of the (dropout-masked) upstream
//     gradient, red.add dx into the word-gradient row (the only table whose rows are mostly distinct), park dx in a
//     shared-memory slab;
//   phase B (thread = 4 columns): walk the parked rows; position rows are summed in registers while consecutive rows
//     carry the same position id (left-aligned batches: all sequences of the group) and flushed with ONE red.add;
//     token-type and item-position rows are accumulated into shared-memory copies of the two tables that the column's
//     owner updates without atomics.
// The shared tables are flushed once per CTA at the end (item rows only if touched).  Global atomics per token drop
// from 4 x 768 to ~1.2 x 768; the former kernel spent its 145 us in the L2 atomic units (12.6 M red.v4 per step,
// 16 384 of them onto each word of the 3 token-type rows).
struct EmbedBwdSmem {
  static constexpr int STAGE_FLOATS = (ROW_THREADS / 32) * 768;
  static size_t bytes(int type_size, int max_item) {
    return (2 * STAGE_FLOATS + static_cast<size_t>(type_size + max_item) * 768) * 4 + 2 * (ROW_THREADS / 32) * 16 +
           static_cast<size_t>(max_item) * 4 + 16;
  }
};

template <int NV4>
__global__ void __launch_bounds__(ROW_THREADS, 1)
embed_ln_bwd_kernel(const EmbedDev a, const __nv_bfloat16* __restrict__ dout, float* d_word, float* d_pos,
                    float* d_type, float* d_item, float* d_gamma, float* d_beta, const uint8_t* __restrict__ active) {
  constexpr int E = NV4 * 128, WARPS = ROW_THREADS / 32;
  extern __shared__ float4 embed_smem[];
  float4* stage = embed_smem;                                   // [2][WARPS][E / 4]
  float4* tab_type = stage + 2 * WARPS * (E / 4);               // [type_size][E / 4]
  float4* tab_item = tab_type + a.type_size * (E / 4);          // [max_item][E / 4]
  int4* keys = reinterpret_cast<int4*>(tab_item + a.max_item * (E / 4));   // [2][WARPS] (pid, tt, ip, valid)
  int* touched = reinterpret_cast<int*>(keys + 2 * WARPS);      // [max_item]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (a.type_size + a.max_item) * (E / 4); i += ROW_THREADS)
    tab_type[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < a.max_item; i += ROW_THREADS) touched[i] = 0;
  __syncthreads();

  const EmbedMap mp = embed_map(a.B, a.Lp, ROW_THREADS / 32);
  const int per = (mp.items + gridDim.x - 1) / gridDim.x;
  const int w0 = blockIdx.x * per, w1 = min(mp.items, w0 + per);
  auto ids_of = [&](int w) -> TokIds {
    TokIds r{a.pad, a.pad, 0, a.pad};
    if (w < w1) {
      int b, l;
      mp.locate(w, warp, b, l);
      if (b < a.B && l < a.Lp) r = embed_load_ids(a, b, l);
    }
    return r;
  };
  float4 dg[NV4], db[NV4];
#pragma unroll
  for (int k = 0; k < NV4; ++k) dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  TokIds cur = ids_of(w0), nxt = ids_of(w0 + 1);
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* word4 = reinterpret_cast<const float4*>(a.word);
  // The two DRAM-latency loads of an item (its random word row, its upstream-gradient row) are issued one item ahead,
  // before phase B and the barrier of the current item, and held in registers.
  auto tok_of = [&](int w) -> int {      // token index of this warp's row of item w, or -1
    if (w >= w1) return -1;
    int b, l;
    mp.locate(w, warp, b, l);
    return (b < a.B && l < a.Lp && row_on(active, b * a.Lp + l)) ? b * a.Lp + l : -1;
  };
  float4 wrow[NV4];
  uint2 drow[NV4];
  {
    const int idc = min(max(cur.id, 0), a.vocab - 1);
    const int t0 = max(tok_of(w0), 0);
#pragma unroll
    for (int k = 0; k < NV4; ++k) {
      wrow[k] = __ldg(word4 + static_cast<size_t>(idc) * (E / 4) + k * 32 + lane);
      drow[k] = reinterpret_cast<const uint2*>(dout + static_cast<size_t>(t0) * E)[k * 32 + lane];
    }
  }
  for (int w = w0; w < w1; ++w) {
    const int buf = (w - w0) & 1;
    const TokIds nn = ids_of(w + 2);
    float4 wnext[NV4];
    uint2 dnext[NV4];
    {
      const int idn = min(max(nxt.id, 0), a.vocab - 1);
      const int tn = max(tok_of(w + 1), 0);
#pragma unroll
      for (int k = 0; k < NV4; ++k) {
        wnext[k] = __ldg(word4 + static_cast<size_t>(idn) * (E / 4) + k * 32 + lane);
        dnext[k] = reinterpret_cast<const uint2*>(dout + static_cast<size_t>(tn) * E)[k * 32 + lane];
      }
    }
    int b, l;
    mp.locate(w, warp, b, l);
    const bool valid = b < a.B && l < a.Lp && row_on(active, b * a.Lp + l);     // warp-uniform
    // ---------------- phase A ----------------
    if (valid) {
      const int t = b * a.Lp + l;
      embed_clamp_ids(a, cur);
      float4 x[NV4];
      float mean, rstd;
      {
        const float4* pw = reinterpret_cast<const float4*>(a.posw + static_cast<size_t>(cur.pid) * E);
        const float4* tw = reinterpret_cast<const float4*>(a.type + static_cast<size_t>(cur.tt) * E);
        const float4* iw = reinterpret_cast<const float4*>(a.item + static_cast<size_t>(cur.ip) * E);
        float sx = 0.f;
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
          const int v = k * 32 + lane;
          const float4 a0 = wrow[k], a1 = __ldg(pw + v), a2 = __ldg(tw + v), a3 = __ldg(iw + v);
          x[k].x = ((a0.x + a1.x) + a2.x) + a3.x;
          x[k].y = ((a0.y + a1.y) + a2.y) + a3.y;
          x[k].z = ((a0.z + a1.z) + a2.z) + a3.z;
          x[k].w = ((a0.w + a1.w) + a2.w) + a3.w;
          sx += (x[k].x + x[k].y) + (x[k].z + x[k].w);
        }
        mean = warp_sum(sx) / E;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NV4; ++k) {
          const float dx = x[k].x - mean, dy = x[k].y - mean, dz = x[k].z - mean, dw = x[k].w - mean;
          q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        rstd = rsqrtf(warp_sum(q) / E + a.eps);
      }
      float4 gy[NV4];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NV4; ++k) {
        const int v = k * 32 + lane;
        const uint2 raw = drow[k];
        float2 lo = unpack_bf16(raw.x), hi = unpack_bf16(raw.y);
        float4 dy = make_float4(lo.x, lo.y, hi.x, hi.y);
        if (a.drop_thresh != 0) {
          const uint32_t keep = dropout_keep4(a.drop_seed, static_cast<uint64_t>(t) * (E / 4) + v, a.drop_thresh);
          dy.x = (keep & 1u) ? dy.x * a.drop_scale : 0.f; dy.y = (keep & 2u) ? dy.y * a.drop_scale : 0.f;
          dy.z = (keep & 4u) ? dy.z * a.drop_scale : 0.f; dy.w = (keep & 8u) ? dy.w * a.drop_scale : 0.f;
        }
        const float4 xh = make_float4((x[k].x - mean) * rstd, (x[k].y - mean) * rstd, (x[k].z - mean) * rstd,
                                      (x[k].w - mean) * rstd);
        x[k] = xh;
        dg[k].x += dy.x * xh.x; dg[k].y += dy.y * xh.y; dg[k].z += dy.z * xh.z; dg[k].w += dy.w * xh.w;
        db[k].x += dy.x; db[k].y += dy.y; db[k].z += dy.z; db[k].w += dy.w;
        const float4 g = __ldg(g4 + v);
        gy[k] = make_float4(dy.x * g.x, dy.y * g.y, dy.z * g.z, dy.w * g.w);
        s1 += (gy[k].x + gy[k].y) + (gy[k].z + gy[k].w);
        s2 += (gy[k].x * xh.x + gy[k].y * xh.y) + (gy[k].z * xh.z + gy[k].w * xh.w);
      }
      s1 = warp_sum(s1) / E;
      s2 = warp_sum(s2) / E;
      float4* srow = stage + (buf * WARPS + warp) * (E / 4);
#pragma unroll
      for (int k = 0; k < NV4; ++k) {
        const int v = k * 32 + lane;
        const float4 dx = make_float4(rstd * (gy[k].x - s1 - x[k].x * s2), rstd * (gy[k].y - s1 - x[k].y * s2),
                                      rstd * (gy[k].z - s1 - x[k].z * s2), rstd * (gy[k].w - s1 - x[k].w * s2));
        // nn.Embedding(padding_idx=pad) never accumulates a gradient into the pad row
        // (ref: recformer/models.py:89,104-106)
        if (d_word && cur.id != a.pad) red_add_v4(d_word + static_cast<size_t>(cur.id) * E + v * 4, dx.x, dx.y, dx.z, dx.w);
        srow[v] = dx;
      }
    }
    if (lane == 0) keys[buf * WARPS + warp] = make_int4(cur.pid, cur.tt, cur.ip, valid ? 1 : 0);
    __syncthreads();
    // ---------------- phase B ----------------
    if (threadIdx.x < E / 4) {
      const int j = threadIdx.x;
      float4 accp = make_float4(0.f, 0.f, 0.f, 0.f);
      int curp = -1;
#pragma unroll 1
      for (int r = 0; r < WARPS; ++r) {
        const int4 key = keys[buf * WARPS + r];
        if (!key.w) continue;
        const float4 v = stage[(buf * WARPS + r) * (E / 4) + j];
        if (d_pos && key.x != a.pad) {
          if (key.x != curp) {
            if (curp >= 0) red_add_v4(d_pos + static_cast<size_t>(curp) * E + j * 4, accp.x, accp.y, accp.z, accp.w);
            accp = make_float4(0.f, 0.f, 0.f, 0.f);
            curp = key.x;
          }
          accp.x += v.x; accp.y += v.y; accp.z += v.z; accp.w += v.w;
        }
        if (d_type) {
          float4 c = tab_type[key.y * (E / 4) + j];
          c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
          tab_type[key.y * (E / 4) + j] = c;
        }
        if (d_item) {
          float4 c = tab_item[key.z * (E / 4) + j];
          c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
          tab_item[key.z * (E / 4) + j] = c;
          if (j == 0) touched[key.z] = 1;
        }
      }
      if (curp >= 0) red_add_v4(d_pos + static_cast<size_t>(curp) * E + j * 4, accp.x, accp.y, accp.z, accp.w);
    }
    // no second barrier: the next item parks its rows in the other slab, and the slab used now is only rewritten
    // two items later, i.e. after the next barrier
#pragma unroll
    for (int k = 0; k < NV4; ++k) { wrow[k] = wnext[k]; drow[k] = dnext[k]; }
    cur = nxt;
    nxt = nn;
  }
  __syncthreads();
  // ---- flush: LayerNorm weight / bias gradients (8 warps -> 1 through the first slab), then the shared tables ----
  float* red = reinterpret_cast<float*>(stage);
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    float* dst = which == 0 ? d_gamma : d_beta;
    if (dst == nullptr) continue;    // uniform
#pragma unroll
    for (int k = 0; k < NV4; ++k)
      reinterpret_cast<float4*>(red + warp * E)[k * 32 + lane] = which == 0 ? dg[k] : db[k];
    __syncthreads();
    if (threadIdx.x < E / 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < WARPS; ++r) {
        const float4 v = reinterpret_cast<const float4*>(red + r * E)[threadIdx.x];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      red_add_v4(dst + threadIdx.x * 4, t.x, t.y, t.z, t.w);
    }
    __syncthreads();
  }
  if (d_type)
    for (int i = threadIdx.x; i < a.type_size * (E / 4); i += ROW_THREADS) {
      const float4 v = tab_type[i];
      red_add_v4(d_type + static_cast<size_t>(i) * 4, v.x, v.y, v.z, v.w);
    }
  if (d_item)
    for (int i = threadIdx.x; i < a.max_item * (E / 4); i += ROW_THREADS) {
      if (!touched[i / (E / 4)]) continue;
      const float4 v = tab_item[i];
      red_add_v4(d_item + static_cast<size_t>(i) * 4, v.x, v.y, v.z, v.w);
    }
}

// ----------------------------------------------------------------------------------------------
// LayerNorm on bf16 rows.  NV8 = E / 256 uint4 (8 x bf16) per lane.
// ----------------------------------------------------------------------------------------------
template <int NV8>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ xr, int lane, float (&v)[NV8][8]) {
#pragma unroll
  for (int k = 0; k < NV8; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(xr + (k * 32 + lane) * 8);
    const float4 b = *reinterpret_cast<const float4*>(xr + (k * 32 + lane) * 8 + 4);
    v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w;
    v[k][4] = b.x; v[k][5] = b.y; v[k][6] = b.z; v[k][7] = b.w;
  }
}

template <int NV8>
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y, float* __restrict__ y32, float* __restrict__ stats, int T,
                     float eps, const uint8_t* __restrict__ active) {
  constexpr int E = NV8 * 256;
  const int lane = threadIdx.x & 31;
  for (int t = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5); t < T; t += gridDim.x * (ROW_THREADS / 32)) {
    if (!row_on(active, t)) continue;
    float v[NV8][8];
    load_row_f32<NV8>(x + static_cast<size_t>(t) * E, lane, v);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV8; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[k][e];
    const float mean = warp_sum(s) / E;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV8; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[k][e] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / E + eps);
    if (lane == 0 && stats) { stats[2 * t] = mean; stats[2 * t + 1] = rstd; }
#pragma unroll
    for (int k = 0; k < NV8; ++k) {
      const int c = (k * 32 + lane) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
      float o[8];
      o[0] = (v[k][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[k][1] - mean) * rstd * g0.y + b0.y;
      o[2] = (v[k][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[k][3] - mean) * rstd * g0.w + b0.w;
      o[4] = (v[k][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[k][5] - mean) * rstd * g1.y + b1.y;
      o[6] = (v[k][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[k][7] - mean) * rstd * g1.w + b1.w;
      if (y) {
        uint4 ov;
        ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]); ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
        reinterpret_cast<uint4*>(y + static_cast<size_t>(t) * E)[k * 32 + lane] = ov;
      }
      if (y32) {
        float* yr = y32 + static_cast<size_t>(t) * E + c;
        *reinterpret_cast<float4*>(yr) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(yr + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

template <int NV8>
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                     const float* __restrict__ stats, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
                     __nv_bfloat16* __restrict__ dx_dropped, float drop_scale, uint32_t drop_thresh,
                     uint64_t drop_seed, float* d_gamma, float* d_beta, float* d_bias, int T,
                     const uint8_t* __restrict__ active, const uint8_t* __restrict__ drop_mask) {
  constexpr int E = NV8 * 256;
  const int lane = threadIdx.x & 31;
  __shared__ float red[ROW_THREADS / 32][E];
  float dg[NV8][8], db[NV8][8], dbias[NV8][8];
#pragma unroll
  for (int k = 0; k < NV8; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) dg[k][e] = db[k][e] = dbias[k][e] = 0.f;
  // The next row's loads are issued before the current row is processed (register double buffer): with one
  // CTA of 8 warps per SM a warp that loads, reduces and stores row after row keeps only 4.5 KB in flight,
  // and the kernel runs at DRAM latency instead of DRAM bandwidth.
  const int t_step = gridDim.x * (ROW_THREADS / 32);
  int t = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5);
  float nx[NV8][8];
  uint4 ndy[NV8];
  uint8_t nmk[NV8];      // saved dropout mask bytes of the row (one per 8 columns), when the GEMM epilogue kept them
#pragma unroll
  for (int k = 0; k < NV8; ++k) nmk[k] = 0xFF;
  float2 nst = make_float2(0.f, 1.f);
  if (t < T && row_on(active, t)) {
    load_row_f32<NV8>(x + static_cast<size_t>(t) * E, lane, nx);
    const uint4* dr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(t) * E);
#pragma unroll
    for (int k = 0; k < NV8; ++k) ndy[k] = dr[k * 32 + lane];
    if (drop_mask != nullptr) {
#pragma unroll
      for (int k = 0; k < NV8; ++k) nmk[k] = drop_mask[static_cast<size_t>(t) * (E / 8) + k * 32 + lane];
    }
    nst = *reinterpret_cast<const float2*>(stats + 2 * t);
  }
  for (; t < T; t += t_step) {
    const float mean = nst.x, rstd = nst.y;
    float xh[NV8][8], gy[NV8][8];
    uint4 cdy[NV8];
    uint8_t cmk[NV8];
#pragma unroll
    for (int k = 0; k < NV8; ++k) {
      cdy[k] = ndy[k];
      cmk[k] = nmk[k];
#pragma unroll
      for (int e = 0; e < 8; ++e) xh[k][e] = nx[k][e];
    }
    if (t + t_step < T && row_on(active, t + t_step)) {
      const int tn = t + t_step;
      load_row_f32<NV8>(x + static_cast<size_t>(tn) * E, lane, nx);
      const uint4* dr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(tn) * E);
#pragma unroll
      for (int k = 0; k < NV8; ++k) ndy[k] = dr[k * 32 + lane];
      if (drop_mask != nullptr) {
#pragma unroll
        for (int k = 0; k < NV8; ++k) nmk[k] = drop_mask[static_cast<size_t>(tn) * (E / 8) + k * 32 + lane];
      }
      nst = *reinterpret_cast<const float2*>(stats + 2 * tn);
    }
    if (!row_on(active, t)) continue;      // padding-only tile: nothing was loaded for this row, nothing is written
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV8; ++k) {
      const int c = (k * 32 + lane) * 8;
      const uint4 draw = cdy[k];
      const float2 d0 = unpack_bf16(draw.x), d1 = unpack_bf16(draw.y), d2 = unpack_bf16(draw.z), d3 = unpack_bf16(draw.w);
      const float dv[8] = {d0.x, d0.y, d1.x, d1.y, d2.x, d2.y, d3.x, d3.y};
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xh[k][e] = (xh[k][e] - mean) * rstd;
        dg[k][e] += dv[e] * xh[k][e];
        db[k][e] += dv[e];
        gy[k][e] = dv[e] * gv[e];
        s1 += gy[k][e];
        s2 += gy[k][e] * xh[k][e];
      }
    }
    s1 = warp_sum(s1) / E;
    s2 = warp_sum(s2) / E;
    uint4* oxr = reinterpret_cast<uint4*>(dx + static_cast<size_t>(t) * E);
#pragma unroll
    for (int k = 0; k < NV8; ++k) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = rstd * (gy[k][e] - s1 - xh[k][e] * s2);
      uint4 ov;
      ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]); ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
      oxr[k * 32 + lane] = ov;
      if (dx_dropped != nullptr) {
        const uint64_t grp = (static_cast<uint64_t>(t) * E + (k * 32 + lane) * 8) >> 3;
        const uint32_t keep = !drop_thresh ? 0xFFu : (drop_mask != nullptr ? cmk[k] : dropout_keep8(drop_seed, grp, drop_thresh));
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = ((keep >> e) & 1u) ? o[e] * drop_scale : 0.f;
        ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]); ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
        reinterpret_cast<uint4*>(dx_dropped + static_cast<size_t>(t) * E)[k * 32 + lane] = ov;
      }
      // column sums of the gradient that reaches the preceding dense layer (its bias gradient)
#pragma unroll
      for (int e = 0; e < 8; ++e) dbias[k][e] += o[e];
    }
  }
  // CTA-level reduction (8 warps -> 1) in shared memory, then one red.add per column per CTA
  const int warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? d_gamma : (which == 1 ? d_beta : d_bias);
    if (dst == nullptr) continue;   // uniform across the CTA
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV8; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e)
        red[warp][(k * 32 + lane) * 8 + e] = which == 0 ? dg[k][e] : (which == 1 ? db[k][e] : dbias[k][e]);
    __syncthreads();
    for (int c = threadIdx.x; c < E; c += ROW_THREADS) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < ROW_THREADS / 32; ++w) t += red[w][c];
      atomicAdd(dst + c, t);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// column sums of a bf16 [T,N] matrix (bias gradients): out[n] += sum_t x[t,n]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* out, int T, int N,
                                                          int ld, int rows_per_cta, const uint8_t* __restrict__ active) {
  // CTA = one 256-column slab (lane -> 8 columns) x rows_per_cta rows (warp w takes rows w, w+8, ...);
  // warps are reduced in shared memory, then one red.add per column per CTA.
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int t0 = blockIdx.y * rows_per_cta;
  const int t1 = min(T, t0 + rows_per_cta);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < N) {
    int t = t0 + warp;
    for (; t + 24 < t1; t += 32) {   // 4 independent loads in flight
      uint4 raw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        raw[q] = row_on(active, t + q * 8) ? *reinterpret_cast<const uint4*>(x + static_cast<size_t>(t + q * 8) * ld + c)
                                           : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 a0 = unpack_bf16(raw[q].x), a1 = unpack_bf16(raw[q].y), a2 = unpack_bf16(raw[q].z), a3 = unpack_bf16(raw[q].w);
        acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
        acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
      }
    }
    for (; t < t1; t += 8) {
      if (!row_on(active, t)) continue;
      const uint4 raw = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(t) * ld + c);
      const float2 a0 = unpack_bf16(raw.x), a1 = unpack_bf16(raw.y), a2 = unpack_bf16(raw.z), a3 = unpack_bf16(raw.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

// ----------------------------------------------------------------------------------------------
// cast + AdamW on the flat parameter buffer
// ----------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

template <bool G_BF16, bool ZERO_G = false>
__global__ void adamw_kernel(float4* __restrict__ p, const void* g_, float4* __restrict__ m,
                             float4* __restrict__ v, uint2* __restrict__ shadow, long long n4, float lr, float b1,
                             float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                             const float* __restrict__ hp) {
  if (hp != nullptr) {   // per-step scalars from device memory (a captured graph cannot take them as arguments)
    lr = __ldg(hp); bc1 = __ldg(hp + 1); bc2_sqrt = __ldg(hp + 2); gscale = __ldg(hp + 3);
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg, mm = m[i], vv = v[i];
    if (G_BF16) {   // gradients as they come off the bf16 all-reduce of data-parallel training
      const uint2 raw = reinterpret_cast<const uint2*>(g_)[i];
      const float2 lo = unpack_bf16(raw.x), hi = unpack_bf16(raw.y);
      gg = make_float4(lo.x, lo.y, hi.x, hi.y);
    } else {
      gg = reinterpret_cast<const float4*>(g_)[i];
      // the step consumed this gradient: leave zeros behind, so that a captured training step needs no separate
      // 600 MB memset of the flat gradient buffer (rf_adamw_step_zero)
      if (ZERO_G) reinterpret_cast<float4*>(const_cast<void*>(g_))[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* pa = reinterpret_cast<float*>(&pp); float* ga = reinterpret_cast<float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm); float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = ga[e] * gscale;
      pa[e] *= (1.f - lr * wd);                      // decoupled weight decay (torch.optim.AdamW)
      ma[e] = b1 * ma[e] + (1.f - b1) * gr;
      va[e] = b2 * va[e] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(va[e]) / bc2_sqrt + eps;
      pa[e] -= (lr / bc1) * (ma[e] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (shadow) shadow[i] = make_uint2(pack_bf16(pa[0], pa[1]), pack_bf16(pa[2], pa[3]));
  }
}

static EmbedDev make_embed_dev(const rf_embed_args* a) {
  EmbedDev d;
  d.ids = a->input_ids; d.tt = a->token_type_ids; d.ip = a->item_position_ids; d.pos = a->pos_ids;
  d.word = a->word_emb; d.posw = a->pos_emb; d.type = a->type_emb; d.item = a->item_emb;
  d.gamma = a->ln_gamma; d.beta = a->ln_beta;
  d.B = a->B; d.L = a->L; d.Lp = a->Lp; d.E = a->E; d.vocab = a->vocab; d.max_pos = a->max_pos;
  d.type_size = a->type_size; d.max_item = a->max_item; d.pad = a->padding_idx; d.eps = a->eps;
  d.drop_thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  d.drop_scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  d.drop_seed = a->drop_seed;
  return d;
}

// grid of the embedding kernels: one CTA per work item (see EmbedMap), capped at `per_sm` CTAs per SM
static int embed_grid(const rf_embed_args* a, int warps, int per_sm) {
  const long long items = embed_map(a->B, a->Lp, warps).items;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  return static_cast<int>(items < cap ? (items > 0 ? items : 1) : cap);
}

static int row_grid(int T) {
  const int need = (T + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32);
  const int cap = sm_count() * 8;
  return need < cap ? (need > 0 ? need : 1) : cap;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_prepare_inputs(const int64_t* input_ids, const int64_t* attention_mask,
                                 const int64_t* global_attention_mask, int B, int L, int Lp, int padding_idx,
                                 int32_t* pos_ids, uint8_t* mask012, int* err_flag, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(input_ids && pos_ids && mask012 && err_flag, "rf_prepare_inputs: null argument");
  RF_REQUIRE(B > 0 && L > 0 && Lp >= L, "rf_prepare_inputs: bad shape B=%d L=%d Lp=%d", B, L, Lp);
  const int warps = 4;
  prepare_inputs_kernel<<<(B + warps - 1) / warps, warps * 32, 0, stream>>>(input_ids, attention_mask,
                                                                          global_attention_mask, B, L, Lp, padding_idx,
                                                                          pos_ids, mask012, err_flag);
  return check_launch("rf_prepare_inputs");
}

extern "C" int rf_row_tile_flags(const uint8_t* mask012, int B, int L, uint8_t* tile_flags, int32_t* qtile_list,
                                 int32_t* n_qtiles, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(mask012 && tile_flags && qtile_list && n_qtiles, "rf_row_tile_flags: null argument");
  RF_REQUIRE(B > 0 && L > 0 && L % 256 == 0, "rf_row_tile_flags: L=%d must be a multiple of 256", L);
  RF_REQUIRE((reinterpret_cast<uintptr_t>(mask012) & 7) == 0, "rf_row_tile_flags: mask must be 8-byte aligned");
  row_tile_flags_kernel<<<1, 256, 0, stream>>>(mask012, B * (L / 256), tile_flags, qtile_list, n_qtiles);
  return check_launch("rf_row_tile_flags");
}

extern "C" int rf_embed_ln_fwd(const rf_embed_args* a, void* out, float* out32, int* err_flag, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && (out || out32), "rf_embed_ln_fwd: null argument");
  RF_REQUIRE(a->E == 768, "rf_embed_ln_fwd: hidden size %d unsupported (768)", a->E);
  RF_REQUIRE(a->B > 0 && a->L > 0 && a->Lp >= a->L, "rf_embed_ln_fwd: bad shape");
  const EmbedDev d = make_embed_dev(a);
  embed_ln_fwd_kernel<6><<<row_grid(a->B * a->Lp), ROW_THREADS, 0, stream>>>(
      d, reinterpret_cast<__nv_bfloat16*>(out), out32, err_flag, active_flags_for(static_cast<long long>(a->B) * a->Lp));
  return check_launch("rf_embed_ln_fwd");
}

extern "C" int rf_embed_ln_bwd(const rf_embed_args* a, const void* dout, float* d_word, float* d_pos, float* d_type,
                               float* d_item, float* d_gamma, float* d_beta, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && dout, "rf_embed_ln_bwd: null argument");
  RF_REQUIRE(a->E == 768, "rf_embed_ln_bwd: hidden size %d unsupported (768)", a->E);
  const EmbedDev d = make_embed_dev(a);
  const size_t smem = EmbedBwdSmem::bytes(a->type_size, a->max_item);
  if (smem <= 227 * 1024) {
    auto kern = embed_ln_bwd_kernel<6>;
    static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
    if (first_use_on_device(&attr_seen))
      RF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    kern<<<embed_grid(a, ROW_THREADS / 32, 1), ROW_THREADS, smem, stream>>>(
        d, reinterpret_cast<const __nv_bfloat16*>(dout), d_word, d_pos, d_type, d_item, d_gamma, d_beta,
        active_flags_for(static_cast<long long>(a->B) * a->Lp));
  } else {
    const int grid = min(row_grid(a->B * a->Lp), sm_count() * 2);
    embed_ln_bwd_atomic_kernel<6><<<grid, ROW_THREADS, 0, stream>>>(
        d, reinterpret_cast<const __nv_bfloat16*>(dout), d_word, d_pos, d_type, d_item, d_gamma, d_beta,
        active_flags_for(static_cast<long long>(a->B) * a->Lp));
  }
  return check_launch("rf_embed_ln_bwd");
}

extern "C" int rf_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* y32,
                                float* stats, int T, int E, float eps, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(x && gamma && beta && (y || y32), "rf_layernorm_fwd: null argument");
  RF_REQUIRE(E == 768, "rf_layernorm_fwd: hidden size %d unsupported (768)", E);
  RF_REQUIRE(T > 0, "rf_layernorm_fwd: T=%d", T);
  layernorm_fwd_kernel<3><<<row_grid(T), ROW_THREADS, 0, stream>>>(x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y),
                                                                  y32, stats, T, eps, active_flags_for(T));
  return check_launch("rf_layernorm_fwd");
}

extern "C" int rf_layernorm_bwd(const void* dy, const float* x, const float* stats, const float* gamma, void* dx,
                                void* dx_dropped, float drop_p, uint64_t drop_seed, float* d_gamma, float* d_beta,
                                float* d_bias, int T, int E, const uint8_t* drop_mask, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(dy && x && stats && gamma && dx, "rf_layernorm_bwd: null argument");
  RF_REQUIRE(E == 768, "rf_layernorm_bwd: hidden size %d unsupported (768)", E);
  const uint32_t thresh = drop_p > 0.f ? static_cast<uint32_t>(drop_p * 65536.0f) : 0u;
  const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const int grid = min(row_grid(T), sm_count());
  layernorm_bwd_kernel<3><<<grid, ROW_THREADS, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), x, stats, gamma,
      reinterpret_cast<__nv_bfloat16*>(dx), reinterpret_cast<__nv_bfloat16*>(dx_dropped), scale, thresh, drop_seed,
      d_gamma, d_beta, d_bias, T, active_flags_for(T), thresh != 0 ? drop_mask : nullptr);
  return check_launch("rf_layernorm_bwd");
}

extern "C" int rf_colsum_bf16(const void* x, float* out, int T, int N, int ld, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(x && out && T > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "rf_colsum_bf16: bad argument");
  const int gx = (N + 255) / 256;
  int gy = (sm_count() * 4) / gx;
  if (gy < 1) gy = 1;
  if (gy > T) gy = T;
  const int rows = (T + gy - 1) / gy;
  gy = (T + rows - 1) / rows;
  colsum_bf16_kernel<<<dim3(gx, gy), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, T, N, ld, rows,
                                                      active_flags_for(T));
  return check_launch("rf_colsum_bf16");
}

extern "C" int rf_cast_f32_to_bf16(const float* x, void* y, long long n, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(x && y && n > 0 && n % 4 == 0, "rf_cast_f32_to_bf16: n=%lld must be a positive multiple of 4", n);
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  cast_f32_bf16_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(reinterpret_cast<const float4*>(x),
                                                                  reinterpret_cast<uint2*>(y), n4);
  return check_launch("rf_cast_f32_to_bf16");
}

extern "C" int rf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow,
                             long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             float grad_scale, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && n % 4 == 0 && step >= 1, "rf_adamw_step: bad argument");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  adamw_kernel<false><<<static_cast<int>(grid), 256, 0, stream>>>(
      reinterpret_cast<float4*>(param), grad, reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), reinterpret_cast<uint2*>(shadow), n4, lr, beta1, beta2, eps, weight_decay,
      bc1, sqrtf(bc2), grad_scale, nullptr);
  return check_launch("rf_adamw_step");
}

extern "C" int rf_adamw_step_bf16grad(float* param, const void* grad_bf16, float* exp_avg, float* exp_avg_sq, void* shadow,
                                      long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                      int step, float grad_scale, const float* hp_dev, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(param && grad_bf16 && exp_avg && exp_avg_sq && n > 0 && n % 4 == 0 && (hp_dev || step >= 1),
             "rf_adamw_step_bf16grad: bad argument");
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  const float bc1 = hp_dev ? 1.f : 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = hp_dev ? 1.f : 1.f - powf(beta2, static_cast<float>(step));
  adamw_kernel<true><<<static_cast<int>(grid), 256, 0, stream>>>(
      reinterpret_cast<float4*>(param), grad_bf16, reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), reinterpret_cast<uint2*>(shadow), n4, lr, beta1, beta2, eps, weight_decay,
      bc1, sqrtf(bc2), grad_scale, hp_dev);
  return check_launch("rf_adamw_step_bf16grad");
}

extern "C" int rf_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow,
                                 long long n, float beta1, float beta2, float eps, float weight_decay,
                                 const float* hp_dev, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(param && grad && exp_avg && exp_avg_sq && hp_dev && n > 0 && n % 4 == 0, "rf_adamw_step_dev: bad argument");
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  adamw_kernel<false><<<static_cast<int>(grid), 256, 0, stream>>>(
      reinterpret_cast<float4*>(param), grad, reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), reinterpret_cast<uint2*>(shadow), n4, 0.f, beta1, beta2, eps, weight_decay,
      1.f, 1.f, 1.f, hp_dev);
  return check_launch("rf_adamw_step_dev");
}

extern "C" int rf_adamw_step_zero(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow, long long n,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                  float grad_scale, const float* hp_dev, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && n % 4 == 0 && (hp_dev || step >= 1),
             "rf_adamw_step_zero: bad argument");
  const long long n4 = n / 4;
  long long grid = (n4 + 255) / 256;
  if (grid > sm_count() * 16) grid = sm_count() * 16;
  const float bc1 = hp_dev ? 1.f : 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = hp_dev ? 1.f : 1.f - powf(beta2, static_cast<float>(step));
  adamw_kernel<false, true><<<static_cast<int>(grid), 256, 0, stream>>>(
      reinterpret_cast<float4*>(param), grad, reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), reinterpret_cast<uint2*>(shadow), n4, lr, beta1, beta2, eps, weight_decay,
      bc1, sqrtf(bc2), grad_scale, hp_dev);
  return check_launch("rf_adamw_step_zero");
}

// ================================================================================================
// Device-side batch assembly of the tokenizer's five-tensor layout (SURVEY.md §8a Spec T;
// ref: recformer/tokenization.py:64-152 `encode` + `padding`) from pre-tokenised items held in HBM
// as CSR arrays.  One CTA per user: the user's items are taken most-recent-first, at most
// max_items of them, rows are truncated to max_tokens and right-padded to the output width L with
// (ids = pad, item position = max_item_pos, token type = 3, masks = 0).  Pure integer work, bit-exact.
// ================================================================================================
namespace rf {

__global__ void __launch_bounds__(256)
assemble_batch_kernel(const int64_t* __restrict__ item_offsets, const int32_t* __restrict__ item_tokens,
                      const uint8_t* __restrict__ item_types, const int64_t* __restrict__ user_offsets,
                      const int64_t* __restrict__ user_items, int L, int max_items, int max_tokens, int bos, int pad,
                      int max_item_pos, int64_t* __restrict__ out_ids, int64_t* __restrict__ out_item_pos,
                      int64_t* __restrict__ out_types, int64_t* __restrict__ out_mask, int64_t* __restrict__ out_global,
                      int32_t* __restrict__ out_len) {
  constexpr int MAXI = 64;                       // max_item_embeddings - 1 <= 63
  __shared__ int s_start[MAXI + 1];              // first token position of the k-th kept item (k = 0 most recent)
  __shared__ long long s_src[MAXI];              // offset of that item's tokens in the CSR arrays
  __shared__ int s_n, s_total;
  const int u = blockIdx.x, tid = threadIdx.x;
  const long long u0 = user_offsets[u], u1 = user_offsets[u + 1];
  if (tid == 0) {
    const int n = static_cast<int>(min(static_cast<long long>(max_items), u1 - u0));
    int pos = 1;                                  // position 0 is <s>
    for (int k = 0; k < n; ++k) {
      const long long item = user_items[u1 - 1 - k];   // reversed: most recent first
      const long long a = item_offsets[item], b = item_offsets[item + 1];
      s_start[k] = pos;
      s_src[k] = a;
      pos += static_cast<int>(b - a);
    }
    s_start[n] = pos;
    s_n = n;
    s_total = min(pos, max_tokens);
  }
  __syncthreads();
  const int n = s_n, total = s_total;
  if (tid == 0 && out_len != nullptr) out_len[u] = total;
  for (int p = tid; p < L; p += 256) {
    long long id = pad, ip = max_item_pos, tt = 3, am = 0, gm = 0;
    if (p < total) {
      am = 1;
      if (p == 0) {
        id = bos; ip = 0; tt = 0; gm = 1;
      } else {
        int lo = 0, hi = n - 1;                   // last k with s_start[k] <= p
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (s_start[mid] <= p) lo = mid; else hi = mid - 1;
        }
        const long long src = s_src[lo] + (p - s_start[lo]);
        id = item_tokens[src];
        tt = item_types[src];
        ip = lo + 1;
      }
    }
    const size_t o = static_cast<size_t>(u) * L + p;
    out_ids[o] = id; out_item_pos[o] = ip; out_types[o] = tt; out_mask[o] = am; out_global[o] = gm;
  }
}

}  // namespace rf

extern "C" int rf_assemble_batch(const int64_t* item_offsets, const int32_t* item_tokens, const uint8_t* item_types,
                                 const int64_t* user_offsets, const int64_t* user_items, int B, int L, int max_items,
                                 int max_tokens, int bos_id, int pad_id, int max_item_pos, int64_t* out_ids,
                                 int64_t* out_item_pos, int64_t* out_types, int64_t* out_mask, int64_t* out_global,
                                 int32_t* out_len, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(item_offsets && item_tokens && item_types && user_offsets && user_items, "rf_assemble_batch: null input");
  RF_REQUIRE(out_ids && out_item_pos && out_types && out_mask && out_global, "rf_assemble_batch: null output");
  RF_REQUIRE(B > 0 && L > 0 && max_items >= 1 && max_items <= 63 && max_tokens >= 1,
             "rf_assemble_batch: bad shape (B=%d L=%d max_items=%d max_tokens=%d)", B, L, max_items, max_tokens);
  rf::assemble_batch_kernel<<<B, 256, 0, stream>>>(item_offsets, item_tokens, item_types, user_offsets, user_items, L,
                                                  max_items, max_tokens, bos_id, pad_id, max_item_pos, out_ids,
                                                  out_item_pos, out_types, out_mask, out_global, out_len);
  return rf::check_launch("rf_assemble_batch");
}
