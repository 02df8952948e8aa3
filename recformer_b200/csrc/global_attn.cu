// Global (CLS) attention row of Longformer (HF:963-1056), re-associated so that the
// key_global / value_global projections are never applied to all L tokens (the reference notes
// "TODO: remove the redundant computation", HF:609):
//
//   q_g   = (Wqg x_cls + bqg) / sqrt(D)                              [E]
//   u_h   = Wkg[h]^T q_g[h]                                          [E]   per head
//   s_hj  = u_h . x_j              (+ q_g[h].bkg[h], constant in j -> cancels in the softmax)
//   p_h   = softmax over valid j   (fp32)
//   m_h   = sum_j p'_hj x_j        (p' = dropout(p))                 [E]
//   out_h = Wvg[h] m_h + bvg[h] * sum_j p'_hj                        [D]
//
// Per sequence this is 2*H*L*E MACs instead of 2*L*E*E: 64x less work than the two full GEMMs, and
// it is memory-bound (x, [B*L, E] bf16, is streamed once per pass), so it runs on CUDA cores:
//   * token-parallel "dots" kernel (s_hj, and dp_hj in backward): the 12 per-sequence vectors sit in
//     registers (3 heads per warp), x rows are streamed with 128-bit loads, 4 tokens x 3 heads are
//     reduced with one 16-value reduce-scatter (16 shuffles instead of 60);
//   * column-sliced "mix" kernels (m_h, and du_h / dx in backward): a CTA owns 32 columns of one
//     sequence and walks all its tokens, so every output element has ONE writer — no atomics;
//   * the 768x768 *_global weight products are batched over the sequences (weights read once).
#include <cuda_bf16.h>
#include <math.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

RF_DEFINE_NONCE_LOADER(global_attn)

namespace rf {

constexpr int GH = 12;    // heads
constexpr int GE = 768;   // hidden
constexpr int GD = 64;    // head dim
constexpr int GBB = 16;   // sequences per batch-block of the column-parallel weight products
constexpr int RBB = 4;    // sequences per batch-block of the row-parallel ones (12 KB of staging: the kernels of this
                          // file stay below ~15 KB of shared memory so that they can share an SM with a band-attention CTA)

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ bool seq_has_global(const uint8_t* mask, int b, int L) {
  return mask[static_cast<size_t>(b) * L] == 2;
}

// keep (scaled) / drop factor of probability (b,h,j): the same counter-based mask in every kernel
__device__ __forceinline__ float drop_factor(uint64_t seed, uint32_t thresh, float scale, uint64_t idx) {
  if (thresh == 0) return 1.0f;
  const uint32_t keep = dropout_keep8(seed, idx >> 3, thresh);
  return ((keep >> (idx & 7)) & 1u) ? scale : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// Weight products, batched over sequences.
//
// rowdot: out[b, r] = W[r, :] . V_b(r)        one warp per weight row r, V staged in shared memory
//   MODE_Q  : V = x_cls[b] (bf16);               q_g = (dot + bqg[r]) / 8                -> qg
//   MODE_OUT: V = m[b, h(r), :];                 ctx[b, 0, r] = dot + bvg[r] * psum[b,h]  (if global)
//   MODE_DQ : V = du[b, h(r), :];                dq = dot / 8 (0 if no global) -> dqf; dbqg[r] += sum_b dq
// grid (E / 8, ceil(B / RBB)), 256 threads (8 rows of the same head per CTA).
// ---------------------------------------------------------------------------------------------
enum { MODE_Q = 0, MODE_OUT = 1, MODE_DQ = 2 };

struct RowdotParams {
  const float* W;             // [E, E]
  const float* bias;          // [E] or null
  const __nv_bfloat16* x;     // MODE_Q: layer input [B*L, E]
  const float* V;             // MODE_OUT / MODE_DQ: [B, H, E]
  const float* psum;          // MODE_OUT: [B, H]
  const uint8_t* mask;
  float* out_f32;             // MODE_Q: qg [B,E]; MODE_DQ: dqf [B,E]
  __nv_bfloat16* ctx;         // MODE_OUT: [B*L, E]
  float* dbias;               // MODE_DQ: dbqg (+=) or null
  int B, L;
};

template <int MODE>
__global__ void __launch_bounds__(256) global_rowdot_kernel(const RowdotParams p) {
  __shared__ __align__(16) float vs[RBB][GE];   // 12 KB
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = blockIdx.x * 8 + warp;
  const int h = (blockIdx.x * 8) / GD;
  const int b0 = blockIdx.y * RBB;
  const int nb = min(RBB, p.B - b0);
  for (int i = tid; i < nb * (GE / 4); i += 256) {
    const int bb = i / (GE / 4), c4 = i % (GE / 4);
    float4 v;
    if (MODE == MODE_Q) {
      const uint2 raw = *reinterpret_cast<const uint2*>(p.x + static_cast<size_t>(b0 + bb) * p.L * GE + c4 * 4);
      const float2 a = unpack_bf16(raw.x), c = unpack_bf16(raw.y);
      v = make_float4(a.x, a.y, c.x, c.y);
    } else {
      v = *reinterpret_cast<const float4*>(p.V + (static_cast<size_t>(b0 + bb) * GH + h) * GE + c4 * 4);
    }
    *reinterpret_cast<float4*>(&vs[bb][c4 * 4]) = v;
  }
  float4 w[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) w[k] = __ldg(reinterpret_cast<const float4*>(p.W + static_cast<size_t>(r) * GE) + k * 32 + lane);
  __syncthreads();
  float mine = 0.f;      // lane bb keeps the result of sequence b0 + bb
  for (int bb = 0; bb < nb; ++bb) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(&vs[bb][(k * 32 + lane) * 4]);
      acc += w[k].x * v.x + w[k].y * v.y + w[k].z * v.z + w[k].w * v.w;
    }
    acc = warp_sum(acc);
    if (lane == bb) mine = acc;
  }
  const int b = b0 + lane;
  const bool have = lane < nb;
  if (MODE == MODE_Q) {
    if (have) p.out_f32[static_cast<size_t>(b) * GE + r] = (mine + p.bias[r]) * 0.125f;
  } else if (MODE == MODE_OUT) {
    if (have && seq_has_global(p.mask, b, p.L))
      p.ctx[static_cast<size_t>(b) * p.L * GE + r] = __float2bfloat16(mine + p.bias[r] * p.psum[b * GH + h]);
  } else {
    float g = 0.f;
    if (have) {
      g = seq_has_global(p.mask, b, p.L) ? mine * 0.125f : 0.f;   // gradient w.r.t. (Wqg x + bqg)
      p.out_f32[static_cast<size_t>(b) * GE + r] = g;
    }
    if (p.dbias != nullptr) {
      g = warp_sum(g);
      if (lane == 0) red_add_f32(p.dbias + r, g);
    }
  }
}

// colmix: out[b, h, e] = sum_d W[h*64 + d, e] * A[b, h*64 + d]     (W^T applied per head)
//   u  = colmix(Wkg, q_g),  dm = colmix(Wvg, dout)
// grid (E / 64, H, ceil(B / 16)), 64 threads; thread = column e, 16 sequences in registers.
// Also, for the dm call, produces dpsum[b,h] = bvg[h] . dout_h, dbvg += dout_h psum, and the fp32 copy
// of dout (zero for sequences without a global token).
struct ColmixParams {
  const float* W;
  const float* A;                 // [B, E] fp32, or null when A is read from dctx
  const __nv_bfloat16* dctx;      // A = row 0 of dctx for sequences with a global token
  const uint8_t* mask;
  float* out;                     // [B, H, E]
  float* out_sum;                 // if non-null: red.add the head's contribution into [B, E] instead (dxcls)
  // dm extras (null otherwise)
  const float* bvg; const float* psum; float* dpsum; float* dbvg; float* doutf;
  int B, L;
};

__global__ void __launch_bounds__(64) global_colmix_kernel(const ColmixParams p) {
  __shared__ __align__(16) float as[GD][GBB];
  const int tid = threadIdx.x;
  const int e = blockIdx.x * 64 + tid, h = blockIdx.y, b0 = blockIdx.z * GBB;
  const int nb = min(GBB, p.B - b0);
  // all 64 weight loads of this thread's column are issued before anything waits on memory
  float w[GD];
  const float* wp = p.W + static_cast<size_t>(h) * GD * GE + e;
#pragma unroll
  for (int d = 0; d < GD; ++d) w[d] = __ldg(wp + static_cast<size_t>(d) * GE);
  for (int i = tid; i < GD * GBB; i += 64) {
    const int d = i / GBB, bb = i % GBB;
    float v = 0.f;
    if (bb < nb) {
      const int b = b0 + bb;
      if (p.A != nullptr) v = p.A[static_cast<size_t>(b) * GE + h * GD + d];
      else if (seq_has_global(p.mask, b, p.L)) v = __bfloat162float(p.dctx[static_cast<size_t>(b) * p.L * GE + h * GD + d]);
    }
    as[d][bb] = v;
  }
  __syncthreads();
  if (p.dpsum != nullptr && blockIdx.x == 0) {   // the extras of the dm call: once per (h, batch block)
    {
      float gsum = 0.f;
      for (int bb = 0; bb < nb; ++bb) {
        const float g = as[tid][bb];
        p.doutf[static_cast<size_t>(b0 + bb) * GE + h * GD + tid] = g;
        gsum += g * p.psum[(b0 + bb) * GH + h];
      }
      if (p.dbvg != nullptr) red_add_f32(p.dbvg + h * GD + tid, gsum);
    }
    if (tid < nb) {
      float t = 0.f;
      for (int d = 0; d < GD; ++d) t += p.bvg[h * GD + d] * as[d][tid];
      p.dpsum[(b0 + tid) * GH + h] = t;
    }
  }
  float acc[GBB];
#pragma unroll
  for (int bb = 0; bb < GBB; ++bb) acc[bb] = 0.f;
#pragma unroll
  for (int d = 0; d < GD; ++d) {
#pragma unroll
    for (int q = 0; q < GBB / 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(&as[d][q * 4]);
      acc[q * 4 + 0] += w[d] * a.x; acc[q * 4 + 1] += w[d] * a.y; acc[q * 4 + 2] += w[d] * a.z; acc[q * 4 + 3] += w[d] * a.w;
    }
  }
#pragma unroll
  for (int bb = 0; bb < GBB; ++bb) {
    if (bb < nb) {
      if (p.out_sum != nullptr) red_add_f32(p.out_sum + static_cast<size_t>(b0 + bb) * GE + e, acc[bb]);
      else p.out[(static_cast<size_t>(b0 + bb) * GH + h) * GE + e] = acc[bb];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dots: per-token dot products against the 12 per-sequence vectors.  grid (ceil(L/64), B), 256 thr.
//   MODE 0: s[b,h,j]  = u_h . x_j                              (-inf for padded keys)
//   MODE 1: dp[b,h,j] = keep_j * scale * (dm_h . x_j + dpsum_h)   (0 for padded keys)
// warp = (head group hg = warp & 3 -> heads 3hg..3hg+2, token half warp >> 2 -> 32 tokens); the
// warp's 3 vectors live in registers (lane holds the 24 columns of its three 16-byte x units).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4 raw, float (&v)[8]) {
  const float2 a0 = unpack_bf16(raw.x), a1 = unpack_bf16(raw.y), a2 = unpack_bf16(raw.z), a3 = unpack_bf16(raw.w);
  v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y; v[4] = a2.x; v[5] = a2.y; v[6] = a3.x; v[7] = a3.y;
}

// 16 values per lane -> lane l ends with the 32-lane sum of value ((l >> 1) & 15)
__device__ __forceinline__ float reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int n = 8 >> step;              // values kept after this step
    const int bit = 16 >> step;           // lane bit that selects the kept half
    const bool hi = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < n) {
        const float send = hi ? v[i] : v[i + n];
        const float keep = hi ? v[i + n] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int MODE>
__global__ void __launch_bounds__(256)
global_dots_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ mask,
                   const float* __restrict__ vecs, const float* __restrict__ dpsum, int L, float drop_scale,
                   uint32_t drop_thresh, uint64_t drop_seed, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hg = warp & 3;
  const int j_begin = blockIdx.x * 64 + (warp >> 2) * 32;
  float u[3][24];
#pragma unroll
  for (int hh = 0; hh < 3; ++hh) {
    const float4* up = reinterpret_cast<const float4*>(vecs + (static_cast<size_t>(b) * GH + hg * 3 + hh) * GE);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 a = __ldg(up + (k * 32 + lane) * 2), c = __ldg(up + (k * 32 + lane) * 2 + 1);
      u[hh][k * 8 + 0] = a.x; u[hh][k * 8 + 1] = a.y; u[hh][k * 8 + 2] = a.z; u[hh][k * 8 + 3] = a.w;
      u[hh][k * 8 + 4] = c.x; u[hh][k * 8 + 5] = c.y; u[hh][k * 8 + 6] = c.z; u[hh][k * 8 + 7] = c.w;
    }
  }
  // lane l receives (after the reduce-scatter) value index (l >> 1) & 15 = t * 3 + hh for t < 4
  const int vi = (lane >> 1) & 15;
  const int my_t = vi / 3, my_h = hg * 3 + vi % 3;
  const float my_dpsum = (MODE == 1 && vi < 12) ? dpsum[b * GH + my_h] : 0.f;
#pragma unroll 1
  for (int j0 = j_begin; j0 < j_begin + 32 && j0 < L; j0 += 4) {
    uint4 raw[4][3];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = min(j0 + t, L - 1);
      const uint4* xr = reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * L + j) * GE);
#pragma unroll
      for (int k = 0; k < 3; ++k) raw[t][k] = xr[k * 32 + lane];
    }
    float v[16];
#pragma unroll
    for (int i = 12; i < 16; ++i) v[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float xv[8];
        unpack8(raw[t][k], xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a0 += xv[i] * u[0][k * 8 + i];
          a1 += xv[i] * u[1][k * 8 + i];
          a2 += xv[i] * u[2][k * 8 + i];
        }
      }
      v[t * 3 + 0] = a0; v[t * 3 + 1] = a1; v[t * 3 + 2] = a2;
    }
    float r = reduce_scatter16(v, lane);
    const int j = j0 + my_t;
    if ((lane & 1) == 0 && vi < 12 && j < L) {
      const bool valid = mask[static_cast<size_t>(b) * L + j] != 0;
      const uint64_t idx = (static_cast<uint64_t>(b) * GH + my_h) * L + j;
      if (MODE == 0) r = valid ? r : -INFINITY;
      else r = valid ? (r + my_dpsum) * drop_factor(drop_seed, drop_thresh, drop_scale, idx) : 0.f;
      out[idx] = r;
    }
  }
}

// ---- softmax over j for each (b,h), in place (p is saved for backward); also writes the dropped,
//      transposed copy pt[b, j, h] = p'_hj (the coefficient layout of the mix kernels) and
//      psum[b,h] = sum_j p'_hj.  grid (B*H), 256 threads ----
__global__ void __launch_bounds__(256)
global_softmax_kernel(float* __restrict__ s, int L, float drop_scale, uint32_t drop_thresh, uint64_t drop_seed,
                      float* __restrict__ pt, float* __restrict__ psum) {
  float* row = s + static_cast<size_t>(blockIdx.x) * L;
  const int b = blockIdx.x / GH, h = blockIdx.x % GH;
  __shared__ float red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY;
  for (int j = tid; j < L; j += 256) m = fmaxf(m, row[j]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  if (m == -INFINITY) m = 0.f;
  float l = 0.f;
  for (int j = tid; j < L; j += 256) l += __expf(row[j] - m);
  l = warp_sum(l);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) l += red[w];
  __syncthreads();
  const float inv = l > 0.f ? 1.f / l : 0.f;
  float ps = 0.f;
  for (int j = tid; j < L; j += 256) {
    const float pr = __expf(row[j] - m) * inv;
    row[j] = pr;
    const float pd = pr * drop_factor(drop_seed, drop_thresh, drop_scale, static_cast<uint64_t>(blockIdx.x) * L + j);
    pt[(static_cast<size_t>(b) * L + j) * 16 + h] = pd;
    ps += pd;
  }
  ps = warp_sum(ps);
  if (lane == 0) red[warp] = ps;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    psum[blockIdx.x] = t;
  }
}

// ds_j = p_j (dp_j - sum_k p_k dp_k), written transposed: dst[b, j, h].  grid (B*H), 256 threads
__global__ void __launch_bounds__(256)
global_bwd_ds_kernel(const float* __restrict__ p, const float* __restrict__ dp, int L, float* __restrict__ dst) {
  const float* pr = p + static_cast<size_t>(blockIdx.x) * L;
  const float* dr = dp + static_cast<size_t>(blockIdx.x) * L;
  const int b = blockIdx.x / GH, h = blockIdx.x % GH;
  __shared__ float red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float t = 0.f;
  for (int j = tid; j < L; j += 256) t += pr[j] * dr[j];
  t = warp_sum(t);
  if (lane == 0) red[warp] = t;
  __syncthreads();
  t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  for (int j = tid; j < L; j += 256) dst[(static_cast<size_t>(b) * L + j) * 16 + h] = pr[j] * (dr[j] - t);
}

// ---------------------------------------------------------------------------------------------
// Token walks over 64-column slices, fed by TMA (MIX_ACC: m and du;  MIX_DX: the dx update).  Work item = (sequence, 64-column slice, token
// block); the persistent CTAs take contiguous runs of items (token block fastest), so a CTA
// changes (sequence, slice) at most a couple of times and flushes its per-(head, column) partial
// sums with red.add into the zeroed output.  A producer warp streams, per stage, the x tile
// [TOK tokens x 64 cols] (3-D tensor map, 128B swizzle), the coefficient rows pt / dst
// [TOK x 16 fp32] (1-D bulk copies) and, in backward, the dx tile; 8 consumer warps (lane = column
// pair, warp = token mod 8) do the FMAs from shared memory.
//   MIX_ACC: out[b,h,c] += sum_j coef[b,j,h] x[b,j,c]      (coef = p' -> m;  coef = ds -> du)
//   MIX_DX : dx[b,j,c]  += sum_h p'_hj dm[b,h,c] + ds_hj u[b,h,c]     (tile = dx itself, no flush)
// ---------------------------------------------------------------------------------------------
enum { MIX_ACC = 0, MIX_DX = 1 };

template <int MODE>
struct MixCfg {
  static constexpr bool DX = MODE == MIX_DX;
  static constexpr int TOK = 32;
  static constexpr int STAGES = 3;
  static constexpr uint32_t X_BYTES = TOK * 128;
  static constexpr uint32_t C_BYTES = TOK * 64;
  // MIX_ACC: x tile + one coefficient block;  MIX_DX: dx tile + two coefficient blocks (p', ds)
  static constexpr uint32_t STAGE_BYTES = X_BYTES + (DX ? 2 : 1) * C_BYTES;
  static constexpr uint32_t RED_BYTES = DX ? 0 : GH * 64 * 4;          // cross-warp sums via shared-memory atomics
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + RED_BYTES + 2 * STAGES * 8 + 1024;
};

struct MixParams {
  const float* pt;    // [B, L, 16] coefficient rows (MIX_ACC: p' or ds;  MIX_DX: p')
  const float* dst;   // MIX_DX: [B, L, 16] ds
  const float* dm;    // MIX_DX: [B, H, E]
  const float* u;     // MIX_DX: [B, H, E]
  float* out;         // MIX_ACC: [B, H, E] (zeroed by the caller): m or du
  __nv_bfloat16* dx;  // MIX_DX
  const float* dxcls; // MIX_DX: [B, E] fp32 gradient of the CLS token's own input, added into row 0
  int B, L, nblk, items;
};

template <int MODE>
__global__ void __launch_bounds__(288, 2)
global_mix_kernel(const __grid_constant__ CUtensorMap tmX, const MixParams p) {
  using C = MixCfg<MODE>;
  constexpr bool DX = C::DX;
  constexpr int TOK = C::TOK, STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  float* red = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);                 // [GH][64] (MIX_ACC)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES + C::RED_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int it0 = static_cast<int>(static_cast<long long>(blockIdx.x) * p.items / gridDim.x);
  const int it1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.items / gridDim.x);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); }
    fence_mbar_init();
  }
  if (!DX) for (int i = tid; i < GH * 64; i += 288) red[i] = 0.f;
  __syncthreads();
  // stage layout: [x (MIX_ACC) or dx (MIX_DX) tile][coefficient rows 0][coefficient rows 1 (MIX_DX)]
  constexpr uint32_t OFF_C0 = C::X_BYTES, OFF_C1 = OFF_C0 + C::C_BYTES;
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      int stage = 0; uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        const int blk = it % p.nblk, slice = (it / p.nblk) % GH, b = it / (p.nblk * GH);
        const int j0 = blk * TOK;
        const int nt = min(TOK, p.L - j0);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * C::STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], C::X_BYTES + (DX ? 2u : 1u) * static_cast<uint32_t>(nt) * 64);
        tma_load_3d(st, &tmX, &full_bar[stage], slice * 64, j0, b);
        bulk_load_1d(st + OFF_C0, p.pt + (static_cast<size_t>(b) * p.L + j0) * 16, nt * 64, &full_bar[stage]);
        if (DX) bulk_load_1d(st + OFF_C1, p.dst + (static_cast<size_t>(b) * p.L + j0) * 16, nt * 64, &full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }
  // ---- consumers ----
  float acc[DX ? 1 : GH][2], dmr[DX ? GH : 1][2], ur[DX ? GH : 1][2];
  if (!DX) {
#pragma unroll
    for (int h = 0; h < GH; ++h) acc[h][0] = acc[h][1] = 0.f;
  }
  int cur_key = -1;    // b * GH + slice of the partial sums held in acc / of the vectors held in dmr, ur
  auto flush = [&](int key) {
    if (DX) return;
    const int b = key / GH, slice = key % GH;
#pragma unroll
    for (int h = 0; h < (DX ? 1 : GH); ++h) {
      atomicAdd(&red[h * 64 + lane * 2], acc[h][0]);
      atomicAdd(&red[h * 64 + lane * 2 + 1], acc[h][1]);
      acc[h][0] = acc[h][1] = 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int i = tid; i < GH * 64; i += 256) {
      red_add_f32(p.out + (static_cast<size_t>(b) * GH + i / 64) * GE + slice * 64 + (i & 63), red[i]);
      red[i] = 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  };
  int stage = 0; uint32_t phase = 0;
  for (int it = it0; it < it1; ++it) {
    const int blk = it % p.nblk, slice = (it / p.nblk) % GH, b = it / (p.nblk * GH);
    const int key = b * GH + slice;
    if (key != cur_key) {
      if (cur_key >= 0) flush(cur_key);
      cur_key = key;
      if (DX) {
#pragma unroll
        for (int h = 0; h < GH; ++h) {
          const size_t o = (static_cast<size_t>(b) * GH + h) * GE + slice * 64 + lane * 2;
          const float2 d = *reinterpret_cast<const float2*>(p.dm + o), w = *reinterpret_cast<const float2*>(p.u + o);
          dmr[h][0] = d.x; dmr[h][1] = d.y; ur[h][0] = w.x; ur[h][1] = w.y;
        }
      }
    }
    const int j0 = blk * TOK;
    const int nt = min(TOK, p.L - j0);
    mbar_wait(&full_bar[stage], phase);
    const uint8_t* st = smem + stage * C::STAGE_BYTES;
#pragma unroll 2
    for (int t = warp; t < nt; t += 8) {
      const uint32_t xo = t * 128 + ((((lane >> 2) ^ (t & 7))) << 4) + (lane & 3) * 4;
      float2 xv = unpack_bf16(*reinterpret_cast<const uint32_t*>(st + xo));      // x (MIX_ACC) or dx (MIX_DX)
      float pv[GH], sv[DX ? GH : 1];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(st + OFF_C0 + t * 64 + q * 16);
        pv[q * 4] = a.x; pv[q * 4 + 1] = a.y; pv[q * 4 + 2] = a.z; pv[q * 4 + 3] = a.w;
        if (DX) {
          const float4 c = *reinterpret_cast<const float4*>(st + OFF_C1 + t * 64 + q * 16);
          sv[q * 4] = c.x; sv[q * 4 + 1] = c.y; sv[q * 4 + 2] = c.z; sv[q * 4 + 3] = c.w;
        }
      }
      if (!DX) {
#pragma unroll
        for (int h = 0; h < GH; ++h) { acc[DX ? 0 : h][0] += pv[h] * xv.x; acc[DX ? 0 : h][1] += pv[h] * xv.y; }
      } else {
#pragma unroll
        for (int h = 0; h < GH; ++h) {
          xv.x += pv[h] * dmr[DX ? h : 0][0] + sv[DX ? h : 0] * ur[DX ? h : 0][0];
          xv.y += pv[h] * dmr[DX ? h : 0][1] + sv[DX ? h : 0] * ur[DX ? h : 0][1];
        }
        if (j0 + t == 0) {   // the CLS token's own input gradient (zero for sequences without a global token)
          const float2 c = *reinterpret_cast<const float2*>(p.dxcls + static_cast<size_t>(b) * GE + slice * 64 + lane * 2);
          xv.x += c.x; xv.y += c.y;
        }
        *reinterpret_cast<uint32_t*>(p.dx + (static_cast<size_t>(b) * p.L + j0 + t) * GE + slice * 64 + lane * 2) =
            pack_bf16(xv.x, xv.y);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == STAGES) { stage = 0; phase ^= 1; }
  }
  if (cur_key >= 0) flush(cur_key);
}

// GBW: batch-reduced outer products into the *_global weight gradients, no atomics:
//   dW[h*64+d, e] += sum_b A[b, h*64+d] * V_b[e]      grid (H, 3 column slices, 3 weights)
//   z=0: Wvg  A = dout,  V = m[b,h,:]     z=1: Wkg  A = q_g,  V = du[b,h,:]
//   z=2: Wqg  A = dq,    V = x_cls[b,:]   (A is zero for sequences without a global token)
__global__ void __launch_bounds__(256)
global_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ mask, int B, int L,
                    const float* __restrict__ doutf, const float* __restrict__ mvec, const float* __restrict__ qg,
                    const float* __restrict__ du, const float* __restrict__ dqf, float* dWvg, float* dWkg,
                    float* dWqg) {
  const int h = blockIdx.x, e = blockIdx.y * 256 + threadIdx.x, z = blockIdx.z;
  const float* A = z == 0 ? doutf : (z == 1 ? qg : dqf);
  float* dW = z == 0 ? dWvg : (z == 1 ? dWkg : dWqg);
  if (dW == nullptr) return;
  __shared__ float as[16][GD];
  float acc[GD];
#pragma unroll
  for (int d = 0; d < GD; ++d) acc[d] = 0.f;
  for (int b0 = 0; b0 < B; b0 += 16) {
    const int nb = min(16, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * GD; i += 256) {
      const int bb = i / GD, d = i % GD;
      const bool on = mask[static_cast<size_t>(b0 + bb) * L] == 2;
      as[bb][d] = on ? A[static_cast<size_t>(b0 + bb) * GE + h * GD + d] : 0.f;
    }
    __syncthreads();
    float v[16];
#pragma unroll
    for (int bb = 0; bb < 16; ++bb) {
      const int b = min(b0 + bb, B - 1);
      if (z == 0) v[bb] = mvec[(static_cast<size_t>(b) * GH + h) * GE + e];
      else if (z == 1) v[bb] = du[(static_cast<size_t>(b) * GH + h) * GE + e];
      else v[bb] = __bfloat162float(x[static_cast<size_t>(b) * L * GE + e]);
    }
#pragma unroll
    for (int bb = 0; bb < 16; ++bb) {
      if (bb < nb) {
#pragma unroll
        for (int d = 0; d < GD; ++d) acc[d] += as[bb][d] * v[bb];
      }
    }
  }
#pragma unroll
  for (int d = 0; d < GD; ++d) dW[static_cast<size_t>(h * GD + d) * GE + e] += acc[d];
}


// ---------------------------------------------------------------------------------------------
// Fused token passes (round 2).  The forward used to stream x twice (scores, then the p'-weighted sum) with a
// softmax kernel in between, the backward three times (dp, softmax backward, du), all on CUDA cores
// (2 x 12 x 768 MACs per token: ~27 M warp instructions per pass, issue-bound at ~55 us).  Both are now ONE pass
// each, with the two contractions on the tensor cores (legacy mma.sync m16n8k16: the N dimension is the 12 heads,
// far too narrow for a tcgen05 tile):
//   CTA = (sequence, 64-token chunk).  The x tile [64 x 768] bf16 is staged once in shared memory (cp.async).
//   phase A  S[64 x 16] = X . V^T, V = the 12 per-sequence vectors (u forward, dm backward; fp32, fed as a
//            bf16 hi + lo pair so the scores keep fp32-grade vectors); 8 warps = 4 token tiles x 2 K halves.
//   softmax  forward: chunk-local max / sums per head, weights e * keep -> P' [16 x 64] bf16; raw scores saved.
//            backward: p recomputed from the saved score and lse; delta_h = dm_h . m_h + dpsum_h psum_h (closed form:
//            no reduction over tokens); ds = p (keep scale (dm.x + dpsum) - delta) -> DS [16 x 64] bf16, pt / dst rows.
//   phase B  [16 x 768] = P' . X (or DS . X): each warp owns 96 columns; B fragments by ldmatrix.trans from the tile.
//   The chunk partials (max, sum, dropped sum, vector) are merged by a tiny kernel (forward: rescaled by the chunk
//   maxima -> m_h, psum_h, lse_h; backward: plain sum -> du_h).
// ---------------------------------------------------------------------------------------------
constexpr int GP_THREADS = 256;
constexpr int GP_TOK = 64;                       // tokens per CTA
constexpr int GP_XLD = GE + 8;                   // padded row (1552 B = 97 x 16 B: conflict-free ldmatrix)
constexpr int GP_PLD = GP_TOK + 8;               // padded P' / DS row (144 B)
constexpr uint32_t GP_OFF_S = GP_TOK * GP_XLD * 2;                    // partial scores [2][64][16] fp32
constexpr uint32_t GP_OFF_P = GP_OFF_S + 2 * GP_TOK * 16 * 4;         // P' / DS tile [16][72] bf16
constexpr uint32_t GP_SMEM = GP_OFF_P + 16 * GP_PLD * 2;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// fp32 pair -> bf16x2 "hi" and the bf16x2 rounding remainder "lo"
__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(x, y);
  const float2 h = unpack_bf16(hi);
  lo = pack_bf16(x - h.x, y - h.y);
}

struct PassParams {
  const __nv_bfloat16* x;     // [B*L, E]
  const uint8_t* mask;
  const float* vecs;          // fwd: u [B,H,E];  bwd: dm [B,H,E]
  float* s;                   // [B,H,L] raw scores (fwd: written, -inf for padded keys;  bwd: read)
  float* cmax; float* csum; float* csumd; float* cm;     // chunk partials, per (b, chunk, h)
  // backward
  const float* dpsum; const float* psum; const float* lse; const float* mvec;
  float* pt; float* dst;      // [B, L, 16] token-major coefficient rows (p' and ds)
  int B, L, C;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed;
};

template <bool BWD>
__global__ void __launch_bounds__(GP_THREADS) global_pass_kernel(const PassParams p) {
  extern __shared__ __align__(16) uint8_t gp_smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(gp_smem);
  float* sS = reinterpret_cast<float*>(gp_smem + GP_OFF_S);
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(gp_smem + GP_OFF_P);
  const int b = blockIdx.y, c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = p.L, j0 = c * GP_TOK;
  const int g = lane >> 2, t4 = lane & 3;

  // ---- stage the x tile (rows past L are clamped: their columns of P' / DS are zero) ----
  for (int i = tid; i < GP_TOK * (GE / 8); i += GP_THREADS) {
    const int r = i / (GE / 8), ch = i % (GE / 8);
    const int j = min(j0 + r, L - 1);
    const __nv_bfloat16* src = p.x + (static_cast<size_t>(b) * L + j) * GE + ch * 8;
    const uint32_t dst = smem_u32(xs + r * GP_XLD + ch * 8);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // backward: delta_h of the heads this warp normalises (needs only saved vectors: overlaps the tile load)
  const int hA = warp, hB = warp + 8;            // softmax stage: warp handles heads warp and warp + 8 (if < 12)
  float deltaA = 0.f, deltaB = 0.f;
  if (BWD) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int h = q ? hB : hA;
      if (h < GH) {
        const float* dmv = p.vecs + (static_cast<size_t>(b) * GH + h) * GE;
        const float* mv = p.mvec + (static_cast<size_t>(b) * GH + h) * GE;
        float d = 0.f;
        for (int e = lane; e < GE; e += 32) d += __ldg(dmv + e) * __ldg(mv + e);
        d = warp_sum(d) + p.dpsum[b * GH + h] * p.psum[b * GH + h];
        if (q) deltaB = d; else deltaA = d;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- phase A: S[64 x 16] = X V^T; warp = (token tile mt, K half kh) ----
  {
    const int mt = warp & 3, kh = warp >> 2;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const uint32_t a_base = smem_u32(xs + (mt * 16 + (lane & 15)) * GP_XLD + (lane >> 4) * 8);
    // B fragment rows: n = g of n-tile 0 -> head g; n-tile 1 -> head 8 + g (heads 12..15 do not exist: zero)
    const float* v0 = p.vecs + (static_cast<size_t>(b) * GH + g) * GE;
    const float* v1 = p.vecs + (static_cast<size_t>(b) * GH + 8 + (g < 4 ? g : 0)) * GE;
    const bool have1 = g < 4;
    // the vector fragments are plain global loads (L1 / L2 hits): fetched TWO k-steps ahead of their use
    auto fetch = [&](int ks, float2 (&f)[4]) {
      const int k0 = kh * 384 + ks * 16;
      f[0] = __ldg(reinterpret_cast<const float2*>(v0 + k0 + 2 * t4));
      f[1] = __ldg(reinterpret_cast<const float2*>(v0 + k0 + 8 + 2 * t4));
      f[2] = have1 ? __ldg(reinterpret_cast<const float2*>(v1 + k0 + 2 * t4)) : make_float2(0.f, 0.f);
      f[3] = have1 ? __ldg(reinterpret_cast<const float2*>(v1 + k0 + 8 + 2 * t4)) : make_float2(0.f, 0.f);
    };
    float2 fa[4], fb[4];
    fetch(0, fa);
    fetch(1, fb);
#pragma unroll 1
    for (int ks = 0; ks < 24; ks += 2) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float2 (&f)[4] = half ? fb : fa;
        const int k0 = kh * 384 + (ks + half) * 16;
        uint32_t a[4];
        ldmatrix_x4(a_base + k0 * 2, a);
        uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
        split_bf16x2(f[0].x, f[0].y, h0, l0); split_bf16x2(f[1].x, f[1].y, h1, l1);
        split_bf16x2(f[2].x, f[2].y, h2, l2); split_bf16x2(f[3].x, f[3].y, h3, l3);
        if (ks + half + 2 < 24) fetch(ks + half + 2, f);       // refill this buffer for two steps later
        mma_bf16_16816(acc[0], a, h0, h1);
        mma_bf16_16816(acc[0], a, l0, l1);
        mma_bf16_16816(acc[1], a, h2, h3);
        mma_bf16_16816(acc[1], a, l2, l3);
      }
    }
    float* o = sS + kh * (GP_TOK * 16);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      *reinterpret_cast<float2*>(o + (mt * 16 + g) * 16 + nt * 8 + 2 * t4) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(o + (mt * 16 + g + 8) * 16 + nt * 8 + 2 * t4) = make_float2(acc[nt][2], acc[nt][3]);
    }
  }
  __syncthreads();

  // ---- softmax stage: warp handles heads hA (and hB if < 12); lane handles tokens lane and lane + 32 ----
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int h = q ? hB : hA;
    if (h >= 16) continue;
    if (h >= GH) {                 // padding rows 12..15 of the P' / DS operand
      sP[h * GP_PLD + lane] = __float2bfloat16(0.f);
      sP[h * GP_PLD + lane + 32] = __float2bfloat16(0.f);
      continue;
    }
    const uint64_t row = (static_cast<uint64_t>(b) * GH + h) * L;
    float sc[2], kf[2];
    bool valid[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int tok = lane + 32 * u, j = j0 + tok;
      valid[u] = j < L && p.mask[static_cast<size_t>(b) * L + (j < L ? j : 0)] != 0;
      sc[u] = sS[tok * 16 + h] + sS[GP_TOK * 16 + tok * 16 + h];
      kf[u] = drop_factor(p.drop_seed, p.drop_thresh, p.drop_scale, row + (j < L ? j : 0));
    }
    if (!BWD) {
      float m = -INFINITY;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        sc[u] = valid[u] ? sc[u] : -INFINITY;
        if (j0 + lane + 32 * u < L) p.s[row + j0 + lane + 32 * u] = sc[u];
        m = fmaxf(m, sc[u]);
      }
      m = warp_max(m);
      float l = 0.f, ld = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float e = valid[u] ? __expf(sc[u] - m) : 0.f;
        const float wd = e * kf[u];
        l += e; ld += wd;
        sP[h * GP_PLD + lane + 32 * u] = __float2bfloat16(wd);
      }
      l = warp_sum(l); ld = warp_sum(ld);
      if (lane == 0) {
        const size_t pi = (static_cast<size_t>(b) * p.C + c) * GH + h;
        p.cmax[pi] = m; p.csum[pi] = l; p.csumd[pi] = ld;
      }
    } else {
      const float dps = p.dpsum[b * GH + h], lse = p.lse[b * GH + h];
      const float delta = q ? deltaB : deltaA;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = j0 + lane + 32 * u;
        float pd = 0.f, ds = 0.f;
        if (valid[u]) {
          const float pr = __expf(p.s[row + j] - lse);
          pd = pr * kf[u];
          ds = pr * (kf[u] * (sc[u] + dps) - delta);
        }
        if (j < L) {
          p.pt[(static_cast<size_t>(b) * L + j) * 16 + h] = pd;
          p.dst[(static_cast<size_t>(b) * L + j) * 16 + h] = ds;
        }
        sP[h * GP_PLD + lane + 32 * u] = __float2bfloat16(ds);
      }
    }
  }
  __syncthreads();

  // ---- phase B: [16 heads x 768] = P' X; warp owns columns [96 warp, 96 warp + 96) ----
  {
    float acc[12][4];
#pragma unroll
    for (int n = 0; n < 12; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
    const uint32_t a_base = smem_u32(sP + (lane & 15) * GP_PLD + (lane >> 4) * 8);
    // ldmatrix.x4.trans over [16 tokens x 16 columns]: lanes 0-15 address token rows of the first 8 columns,
    // lanes 16-31 the same rows of the next 8 -> {b0, b1} of n-tile 2i and {b0, b1} of n-tile 2i + 1
    const uint32_t b_base = smem_u32(xs + (lane & 15) * GP_XLD + warp * 96 + (lane >> 4) * 8);
#pragma unroll
    for (int ks = 0; ks < GP_TOK / 16; ++ks) {
      uint32_t a[4];
      ldmatrix_x4(a_base + ks * 32, a);
#pragma unroll
      for (int np = 0; np < 6; ++np) {
        uint32_t bb[4];
        ldmatrix_x4_trans(b_base + (ks * 16 * GP_XLD + np * 16) * 2, bb);
        mma_bf16_16816(acc[2 * np], a, bb[0], bb[1]);
        mma_bf16_16816(acc[2 * np + 1], a, bb[2], bb[3]);
      }
    }
    const size_t pbase = (static_cast<size_t>(b) * p.C + c) * GH;
#pragma unroll
    for (int n = 0; n < 12; ++n) {
      const int col = warp * 96 + n * 8 + 2 * t4;
      *reinterpret_cast<float2*>(p.cm + (pbase + g) * GE + col) = make_float2(acc[n][0], acc[n][1]);
      if (g < 4) *reinterpret_cast<float2*>(p.cm + (pbase + 8 + g) * GE + col) = make_float2(acc[n][2], acc[n][3]);
    }
  }
}

// backward chunk partials -> du[b,h,:] = sum over chunks.  grid (H, B), 256 threads (3 columns each)
__global__ void __launch_bounds__(256)
global_sum_partials_kernel(const float* __restrict__ cm, int C, float* __restrict__ out) {
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* v = cm + ((static_cast<size_t>(b) * C + c) * GH + h) * GE;
    a0 += v[tid]; a1 += v[256 + tid]; a2 += v[512 + tid];
  }
  float* o = out + (static_cast<size_t>(b) * GH + h) * GE;
  o[tid] = a0; o[256 + tid] = a1; o[512 + tid] = a2;
}

// chunk partials -> m_h = sum_j p'_hj x_j, psum_h = sum_j p'_hj, lse_h.  grid (H, B), 256 threads (3 columns each)
__global__ void __launch_bounds__(256)
global_merge_kernel(const float* __restrict__ cmax, const float* __restrict__ csum, const float* __restrict__ csumd,
                    const float* __restrict__ cm, int C, float* __restrict__ mvec, float* __restrict__ psum,
                    float* __restrict__ lse) {
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const size_t base = (static_cast<size_t>(b) * C) * GH + h;
  __shared__ float s_r[64], s_z[3];
  if (tid < 32) {                      // chunk rescale factors first (C <= 64), so that the column loop below is
    float m = -INFINITY;               // a stream of independent loads
    for (int c = tid; c < C; c += 32) m = fmaxf(m, cmax[base + static_cast<size_t>(c) * GH]);
    m = warp_max(m);
    float z = 0.f, zd = 0.f;
    for (int c = tid; c < C; c += 32) {
      const size_t pi = base + static_cast<size_t>(c) * GH;
      const float mc = cmax[pi];
      const float r = (mc > -INFINITY) ? __expf(mc - m) : 0.f;
      s_r[c] = r;
      z += csum[pi] * r; zd += csumd[pi] * r;
    }
    z = warp_sum(z); zd = warp_sum(zd);
    if (tid == 0) { s_z[0] = z; s_z[1] = zd; s_z[2] = m; }
  }
  __syncthreads();
  const float Z = s_z[0], Zd = s_z[1], M = s_z[2];
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* v = cm + (base + static_cast<size_t>(c) * GH) * GE;
    const float r = s_r[c];
    a0 += v[tid] * r; a1 += v[256 + tid] * r; a2 += v[512 + tid] * r;
  }
  const float inv = Z > 0.f ? 1.f / Z : 0.f;
  float* o = mvec + (static_cast<size_t>(b) * GH + h) * GE;
  o[tid] = a0 * inv; o[256 + tid] = a1 * inv; o[512 + tid] = a2 * inv;
  if (tid == 0) {
    psum[b * GH + h] = Zd * inv;
    lse[b * GH + h] = Z > 0.f ? M + logf(Z) : 0.f;
  }
}

}  // namespace rf

using namespace rf;

template <int MODE>
static int launch_mix(const __nv_bfloat16* tile_src, MixParams& p, cudaStream_t stream, const char* what) {
  using C = MixCfg<MODE>;
  auto kern = global_mix_kernel<MODE>;
  static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
  if (first_use_on_device(&attr_seen)) {
    RF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  }
  const uint64_t B = p.B, L = p.L;
  const CUtensorMap* tmx = get_tmap_3d(tile_src, B, L, GE, GE, L * GE, C::TOK);
  if (!tmx) return RF_ERR_CUDA;
  p.nblk = (p.L + C::TOK - 1) / C::TOK;
  p.items = p.B * GH * p.nblk;
  const int grid = p.items < 2 * sm_count() ? p.items : 2 * sm_count();
  kern<<<grid, 288, C::SMEM, stream>>>(*tmx, p);
  return check_launch(what);
}

extern "C" long long rf_global_attn_fwd_ws_bytes(int B, int L, int H) {
  const long long C = (L + GP_TOK - 1) / GP_TOK;
  return 4ll * B * C * H * (3 + static_cast<long long>(H) * GD) + 256;
}

extern "C" int rf_global_attn_fwd(const rf_global_args* a, void* ctx, float* qg, float* u, float* p, float* pt,
                                  float* mvec, float* psum, float* ws, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && ctx && qg && u && p && pt && mvec && psum && ws, "rf_global_attn_fwd: null argument");
  RF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "rf_global_attn_fwd: ws must be 16-byte aligned");
  RF_REQUIRE(a->H == GH && a->D == GD, "rf_global_attn_fwd: only H=12, D=64 supported (got %d, %d)", a->H, a->D);
  RF_REQUIRE(a->B > 0 && a->L > 0 && a->L <= 64 * GP_TOK, "rf_global_attn_fwd: bad shape (L <= %d)", 64 * GP_TOK);
  RF_REQUIRE((reinterpret_cast<uintptr_t>(pt) & 15) == 0, "rf_global_attn_fwd: pt must be 16-byte aligned");
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  const int B = a->B, L = a->L;
  const int bblocks = (B + GBB - 1) / GBB;
  const uint32_t thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  const float scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  int rc;
  {
    RowdotParams r{};
    r.W = a->Wqg; r.bias = a->bqg; r.x = x; r.mask = a->mask012; r.out_f32 = qg; r.B = B; r.L = L;
    global_rowdot_kernel<MODE_Q><<<dim3(GE / 8, (B + RBB - 1) / RBB), 256, 0, stream>>>(r);
    if ((rc = check_launch("rf_global_attn_fwd/q"))) return rc;
  }
  {
    ColmixParams c{};
    c.W = a->Wkg; c.A = qg; c.mask = a->mask012; c.out = u; c.B = B; c.L = L;
    global_colmix_kernel<<<dim3(GE / 64, GH, bblocks), 64, 0, stream>>>(c);
    if ((rc = check_launch("rf_global_attn_fwd/u"))) return rc;
  }
  {
    // one pass over x: scores (saved), online softmax, p'-weighted sum -> chunk partials in the workspace
    const int C = (L + GP_TOK - 1) / GP_TOK;
    float* w = ws;
    PassParams q{};
    q.x = x; q.mask = a->mask012; q.vecs = u; q.s = p; q.B = B; q.L = L; q.C = C;
    q.cmax = w; q.csum = q.cmax + static_cast<size_t>(B) * C * GH; q.csumd = q.csum + static_cast<size_t>(B) * C * GH;
    q.cm = q.csumd + static_cast<size_t>(B) * C * GH;
    q.drop_scale = scale; q.drop_thresh = thresh; q.drop_seed = a->drop_seed;
    static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
    if (first_use_on_device(&attr_seen)) {
      RF_CUDA(cudaFuncSetAttribute(global_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM));
    }
    global_pass_kernel<false><<<dim3(C, B), GP_THREADS, GP_SMEM, stream>>>(q);
    if ((rc = check_launch("rf_global_attn_fwd/pass"))) return rc;
    global_merge_kernel<<<dim3(GH, B), 256, 0, stream>>>(q.cmax, q.csum, q.csumd, q.cm, C, mvec, psum,
                                                        psum + static_cast<size_t>(B) * GH);
    if ((rc = check_launch("rf_global_attn_fwd/merge"))) return rc;
  }
  {
    RowdotParams r{};
    r.W = a->Wvg; r.bias = a->bvg; r.V = mvec; r.psum = psum; r.mask = a->mask012;
    r.ctx = reinterpret_cast<__nv_bfloat16*>(ctx); r.B = B; r.L = L;
    global_rowdot_kernel<MODE_OUT><<<dim3(GE / 8, (B + RBB - 1) / RBB), 256, 0, stream>>>(r);
  }
  return check_launch("rf_global_attn_fwd/out");
}

extern "C" long long rf_global_attn_bwd_ws_bytes(int B, int L, int H) {
  const long long E = static_cast<long long>(H) * GD;
  const long long C = (L + 63) / 64;       // chunk partials of the fused token pass (GP_TOK = 64)
  return 4ll * (2ll * B * H * E + B * H + static_cast<long long>(B) * H * L + 3ll * B * E + 16ll * B * L + B * C * H * E) + 256;
}

namespace {
struct BwdWs {
  float *dm, *du, *dst, *dp, *dpsum, *doutf, *dqf, *dxcls, *part;
  BwdWs(float* ws, int B, int L) {
    dm = ws;
    du = dm + static_cast<size_t>(B) * GH * GE;
    dst = du + static_cast<size_t>(B) * GH * GE;          // [B, L, 16] (64-byte rows: bulk-copy source)
    dp = dst + static_cast<size_t>(B) * L * 16;
    dpsum = dp + static_cast<size_t>(B) * GH * L;
    doutf = dpsum + static_cast<size_t>(B) * GH;
    dqf = doutf + static_cast<size_t>(B) * GE;
    dxcls = dqf + static_cast<size_t>(B) * GE;
    part = dxcls + static_cast<size_t>(B) * GE;           // [B, C, H, E] chunk partials of du
  }
};

// Packs the token-gradient terms of the CLS row as two bf16 GEMM operands (see rf_global_attn_bwd_xk):
//   cf [B*L, 64]:  cols 0..11 = p'_h, 16..27 = ds_h, 28 = [token is the CLS], others 0
//   dmu[B*64, E]:  rows 0..11 = dm_h, 16..27 = u_h,  28 = dx_cls,             others 0
// One thread per 8 output elements (one 16-byte store).
__global__ void __launch_bounds__(256)
global_pack_xk_kernel(const float* __restrict__ pt, const float* __restrict__ dst, const float* __restrict__ dm,
                      const float* __restrict__ u, const float* __restrict__ dxcls, int B, int L,
                      __nv_bfloat16* __restrict__ cf, __nv_bfloat16* __restrict__ dmu) {
  const long long n_cf = static_cast<long long>(B) * L * 8;
  const long long n_dmu = static_cast<long long>(B) * 64 * (GE / 8);
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 0.f;
  __nv_bfloat16* out;
  if (i < n_cf) {
    const long long t = i >> 3;
    const int g = static_cast<int>(i & 7);
    if (g < 4) {
      const float* src = ((g < 2) ? pt : dst) + t * 16 + (g & 1) * 8;
      const float4 a = *reinterpret_cast<const float4*>(src);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      if ((g & 1) == 0) {
        const float4 b = *reinterpret_cast<const float4*>(src + 4);
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else if (g == 3) {
        v[4] = (t % L == 0) ? 1.f : 0.f;
      }
    }
    out = cf + i * 8;
  } else if (i < n_cf + n_dmu) {
    const long long r = i - n_cf;
    const int c8 = static_cast<int>(r % (GE / 8));
    const int k = static_cast<int>((r / (GE / 8)) & 63);
    const int b = static_cast<int>(r / (64 * (GE / 8)));
    const float* src = nullptr;
    if (k < GH) src = dm + (static_cast<size_t>(b) * GH + k) * GE;
    else if (k >= 16 && k < 16 + GH) src = u + (static_cast<size_t>(b) * GH + (k - 16)) * GE;
    else if (k == 28) src = dxcls + static_cast<size_t>(b) * GE;
    if (src != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(src + c8 * 8);
      const float4 c = *reinterpret_cast<const float4*>(src + c8 * 8 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    }
    out = dmu + r * 8;
  } else {
    return;
  }
  uint4 o;
  o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(out) = o;
}
}  // namespace

extern "C" int rf_global_attn_bwd_xk(const rf_global_args* a, const float* u, const float* pt, const float* ws,
                                     void* cf, void* dmu, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && u && pt && ws && cf && dmu, "rf_global_attn_bwd_xk: null argument");
  RF_REQUIRE(a->H == GH && a->D == GD, "rf_global_attn_bwd_xk: only H=12, D=64 supported (got %d, %d)", a->H, a->D);
  RF_REQUIRE(((reinterpret_cast<uintptr_t>(cf) | reinterpret_cast<uintptr_t>(dmu) | reinterpret_cast<uintptr_t>(u)) & 15) == 0,
             "rf_global_attn_bwd_xk: cf, dmu and u must be 16-byte aligned");
  const int B = a->B, L = a->L;
  const BwdWs w(const_cast<float*>(ws), B, L);
  const long long n = static_cast<long long>(B) * L * 8 + static_cast<long long>(B) * 64 * (GE / 8);
  global_pack_xk_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      pt, w.dst, w.dm, u, w.dxcls, B, L, reinterpret_cast<__nv_bfloat16*>(cf),
      reinterpret_cast<__nv_bfloat16*>(dmu));
  return check_launch("rf_global_attn_bwd_xk/pack");
}

// Part B of the backward: the gradient the CLS row sends to every token, added into dx; needs the
// workspace left by rf_global_attn_bwd (dm, ds, dq).  Separate so that part A can run on a side
// stream while dx is still being produced by the QKV dgrad GEMM.
extern "C" int rf_global_attn_bwd_dx(const rf_global_args* a, const float* u, const float* pt, void* dx,
                                     const float* ws, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && u && pt && dx && ws, "rf_global_attn_bwd_dx: null argument");
  RF_REQUIRE(a->H == GH && a->D == GD, "rf_global_attn_bwd_dx: only H=12, D=64 supported (got %d, %d)", a->H, a->D);
  const int B = a->B, L = a->L;
  const BwdWs w(const_cast<float*>(ws), B, L);
  MixParams m{};
  m.pt = pt; m.dst = w.dst; m.dm = w.dm; m.u = u; m.dx = reinterpret_cast<__nv_bfloat16*>(dx); m.dxcls = w.dxcls;
  m.B = B; m.L = L;
  return launch_mix<MIX_DX>(m.dx, m, stream, "rf_global_attn_bwd_dx/dx");
}

extern "C" int rf_global_attn_bwd(const rf_global_args* a, const void* dctx, const float* qg, const float* u,
                                  const float* p, const float* pt, const float* mvec, const float* psum, void* dx,
                                  float* dWqg, float* dbqg, float* dWkg, float* dWvg, float* dbvg, float* ws,
                                  rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && dctx && qg && u && p && pt && mvec && psum && ws, "rf_global_attn_bwd: null argument");
  RF_REQUIRE(a->H == GH && a->D == GD, "rf_global_attn_bwd: only H=12, D=64 supported (got %d, %d)", a->H, a->D);
  const int B = a->B, L = a->L;
  const int bblocks = (B + GBB - 1) / GBB;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  const BwdWs w(ws, B, L);
  RF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(pt) & 15) == 0,
             "rf_global_attn_bwd: ws and pt must be 16-byte aligned");
  const uint32_t thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  const float scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  int rc;
  {
    // dm[b,h,:] = Wvg[h]^T dout_h; dpsum = bvg[h].dout_h; dbvg += dout_h psum; fp32 copy of dout
    ColmixParams c{};
    c.W = a->Wvg; c.A = nullptr; c.dctx = reinterpret_cast<const __nv_bfloat16*>(dctx); c.mask = a->mask012; c.out = w.dm;
    c.bvg = a->bvg; c.psum = psum; c.dpsum = w.dpsum; c.dbvg = dbvg; c.doutf = w.doutf; c.B = B; c.L = L;
    global_colmix_kernel<<<dim3(GE / 64, GH, bblocks), 64, 0, stream>>>(c);
    if ((rc = check_launch("rf_global_attn_bwd/dm"))) return rc;
  }
  {
    // one pass over x: p recomputed from the saved scores / lse, softmax backward with the closed-form delta,
    // du += ds x, and the coefficient rows pt (= p') / dst (= ds) of the dx update
    PassParams q{};
    q.x = x; q.mask = a->mask012; q.vecs = w.dm; q.s = const_cast<float*>(p); q.B = B; q.L = L;
    q.C = (L + GP_TOK - 1) / GP_TOK;
    q.dpsum = w.dpsum; q.psum = psum; q.lse = psum + static_cast<size_t>(B) * GH; q.mvec = mvec;
    q.pt = const_cast<float*>(pt); q.dst = w.dst; q.cm = w.part;
    q.drop_scale = scale; q.drop_thresh = thresh; q.drop_seed = a->drop_seed;
    static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
    if (first_use_on_device(&attr_seen)) {
      RF_CUDA(cudaFuncSetAttribute(global_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM));
    }
    global_pass_kernel<true><<<dim3(q.C, B), GP_THREADS, GP_SMEM, stream>>>(q);
    if ((rc = check_launch("rf_global_attn_bwd/pass"))) return rc;
    global_sum_partials_kernel<<<dim3(GH, B), 256, 0, stream>>>(w.part, q.C, w.du);
    if ((rc = check_launch("rf_global_attn_bwd/du"))) return rc;
  }
  {
    // dq = Wkg[h] du_h / 8 (fp32 copy kept), dbqg += dq
    RowdotParams r{};
    r.W = a->Wkg; r.V = w.du; r.mask = a->mask012; r.out_f32 = w.dqf; r.dbias = dbqg; r.B = B; r.L = L;
    global_rowdot_kernel<MODE_DQ><<<dim3(GE / 8, (B + RBB - 1) / RBB), 256, 0, stream>>>(r);
    if ((rc = check_launch("rf_global_attn_bwd/dq"))) return rc;
  }
  {
    // dxcls[b,:] = Wqg^T dq_b (the CLS token's own input gradient; added into dx row 0 by the dx-update kernel)
    RF_CUDA(cudaMemsetAsync(w.dxcls, 0, static_cast<size_t>(B) * GE * sizeof(float), stream));
    ColmixParams c{};
    c.W = a->Wqg; c.A = w.dqf; c.mask = a->mask012; c.out_sum = w.dxcls; c.B = B; c.L = L;
    global_colmix_kernel<<<dim3(GE / 64, GH, bblocks), 64, 0, stream>>>(c);
    if ((rc = check_launch("rf_global_attn_bwd/dxcls"))) return rc;
  }
  // the three weight-gradient outer products feed only the optimiser: with dWqg == dWkg == dWvg == NULL they are left
  // to rf_global_attn_bwd_wgrad (same ws), which the engine launches AFTER the token-gradient operands the main stream
  // waits for (rf_global_attn_bwd_xk), so the 16-40 us they take are off the backward's critical chain
  if (dWvg != nullptr || dWkg != nullptr || dWqg != nullptr) {
    RF_REQUIRE(dWvg && dWkg && dWqg, "rf_global_attn_bwd: pass all three *_global weight gradients or none");
    global_wgrad_kernel<<<dim3(GH, 3, 3), 256, 0, stream>>>(x, a->mask012, B, L, w.doutf, mvec, qg, w.du, w.dqf, dWvg,
                                                           dWkg, dWqg);
    if ((rc = check_launch("rf_global_attn_bwd/wgrad"))) return rc;
  }
  if (dx != nullptr) return rf_global_attn_bwd_dx(a, u, pt, dx, ws, stream_);
  return RF_OK;
}

extern "C" int rf_global_attn_bwd_wgrad(const rf_global_args* a, const float* qg, const float* mvec, const float* ws,
                                        float* dWqg, float* dWkg, float* dWvg, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && qg && mvec && ws && dWqg && dWkg && dWvg, "rf_global_attn_bwd_wgrad: null argument");
  RF_REQUIRE(a->H == GH && a->D == GD, "rf_global_attn_bwd_wgrad: only H=12, D=64 supported (got %d, %d)", a->H, a->D);
  const BwdWs w(const_cast<float*>(ws), a->B, a->L);
  global_wgrad_kernel<<<dim3(GH, 3, 3), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a->x), a->mask012, a->B, a->L,
                                                         w.doutf, mvec, qg, w.du, w.dqf, dWvg, dWkg, dWqg);
  return check_launch("rf_global_attn_bwd_wgrad");
}
