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
// Per sequence this is 2*H*L*E MACs instead of 2*L*E*E: 64x less work than the two full GEMMs,
// so plain CUDA-core kernels (coalesced, warp-shuffle reductions) are sufficient; they are
// bound by reading x (L x E bf16) twice from L2/HBM.
#include <cuda_bf16.h>
#include <math.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

namespace rf {

constexpr int GH = 12;    // heads
constexpr int GE = 768;   // hidden
constexpr int GD = 64;    // head dim
constexpr int TOK_PER_CTA = 64;

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// ---- G1: q_g and u_h; also zeroes the m / psum accumulators.  grid (H, B), 256 threads ----
__global__ void __launch_bounds__(256)
global_qu_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ mask, const float* __restrict__ Wqg,
                 const float* __restrict__ bqg, const float* __restrict__ Wkg, int L, float* __restrict__ qg,
                 float* __restrict__ u, float* __restrict__ mvec, float* __restrict__ psum) {
  const int h = blockIdx.x, b = blockIdx.y;
  __shared__ float xs[GE];
  __shared__ float qs[GD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* xc = x + static_cast<size_t>(b) * L * GE;   // row 0 of the sequence
  for (int e = tid; e < GE; e += 256) {
    xs[e] = __bfloat162float(xc[e]);
    mvec[(static_cast<size_t>(b) * GH + h) * GE + e] = 0.f;
  }
  if (tid == 0) psum[b * GH + h] = 0.f;
  __syncthreads();
  for (int d = warp; d < GD; d += 8) {
    const float* wr = Wqg + static_cast<size_t>(h * GD + d) * GE;
    float acc = 0.f;
    for (int e = lane; e < GE; e += 32) acc += wr[e] * xs[e];
    acc = warp_sum(acc);
    if (lane == 0) {
      const float q = (acc + bqg[h * GD + d]) * 0.125f;
      qs[d] = q;
      qg[static_cast<size_t>(b) * GE + h * GD + d] = q;
    }
  }
  __syncthreads();
  for (int e = tid; e < GE; e += 256) {
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < GD; ++d) acc += Wkg[static_cast<size_t>(h * GD + d) * GE + e] * qs[d];
    u[(static_cast<size_t>(b) * GH + h) * GE + e] = acc;
  }
}

// ---- G2: raw scores s[b,h,j] = u_h . x_j  (-inf for padded keys).  grid (L/64, B) ----
__global__ void __launch_bounds__(256)
global_scores_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ mask,
                     const float* __restrict__ u, int L, float* __restrict__ s) {
  const int b = blockIdx.y, j0 = blockIdx.x * TOK_PER_CTA;
  extern __shared__ float us[];   // [GH][GE]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ub = u + static_cast<size_t>(b) * GH * GE;
  for (int i = tid; i < GH * GE; i += 256) us[i] = ub[i];
  __syncthreads();
  for (int jj = warp; jj < TOK_PER_CTA; jj += 8) {
    const int j = j0 + jj;
    if (j >= L) break;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * L + j) * GE);
    float xv[24];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint4 raw = xr[k * 32 + lane];
      const float2 a0 = unpack_bf16(raw.x), a1 = unpack_bf16(raw.y), a2 = unpack_bf16(raw.z), a3 = unpack_bf16(raw.w);
      xv[k * 8 + 0] = a0.x; xv[k * 8 + 1] = a0.y; xv[k * 8 + 2] = a1.x; xv[k * 8 + 3] = a1.y;
      xv[k * 8 + 4] = a2.x; xv[k * 8 + 5] = a2.y; xv[k * 8 + 6] = a3.x; xv[k * 8 + 7] = a3.y;
    }
    const bool valid = mask[static_cast<size_t>(b) * L + j] != 0;
#pragma unroll
    for (int h = 0; h < GH; ++h) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* up = us + h * GE + (k * 32 + lane) * 8;
        const float4 u0 = *reinterpret_cast<const float4*>(up), u1 = *reinterpret_cast<const float4*>(up + 4);
        acc += xv[k * 8 + 0] * u0.x + xv[k * 8 + 1] * u0.y + xv[k * 8 + 2] * u0.z + xv[k * 8 + 3] * u0.w +
               xv[k * 8 + 4] * u1.x + xv[k * 8 + 5] * u1.y + xv[k * 8 + 6] * u1.z + xv[k * 8 + 7] * u1.w;
      }
      acc = warp_sum(acc);
      if (lane == 0) s[(static_cast<size_t>(b) * GH + h) * L + j] = valid ? acc : -INFINITY;
    }
  }
}

// ---- G2b: in-place softmax over j for each (b,h).  grid (B*H), 256 threads ----
__global__ void __launch_bounds__(256) global_softmax_kernel(float* __restrict__ s, int L) {
  float* row = s + static_cast<size_t>(blockIdx.x) * L;
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
  const float inv = l > 0.f ? 1.f / l : 0.f;
  for (int j = tid; j < L; j += 256) row[j] = __expf(row[j] - m) * inv;
}

// ---- G3: m[b,h,:] += sum_{j in chunk} p'[b,h,j] x[b,j,:].  grid (L/64, B), 256 threads ----
__global__ void __launch_bounds__(256)
global_mix_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ p, int L, float drop_scale,
                  uint32_t drop_thresh, uint64_t drop_seed, float* __restrict__ mvec, float* __restrict__ psum) {
  const int b = blockIdx.y, j0 = blockIdx.x * TOK_PER_CTA;
  __shared__ float ps[GH][TOK_PER_CTA];
  const int tid = threadIdx.x;
  for (int i = tid; i < GH * TOK_PER_CTA; i += 256) {
    const int h = i / TOK_PER_CTA, jj = i % TOK_PER_CTA;
    const int j = j0 + jj;
    float v = 0.f;
    if (j < L) {
      v = p[(static_cast<size_t>(b) * GH + h) * L + j];
      if (drop_thresh != 0) {
        const uint64_t idx = (static_cast<uint64_t>(b) * GH + h) * L + j;
        const uint32_t keep = dropout_keep8(drop_seed, idx >> 3, drop_thresh);
        v = ((keep >> (idx & 7)) & 1u) ? v * drop_scale : 0.f;
      }
    }
    ps[h][jj] = v;
  }
  __syncthreads();
  if (tid < GH) {   // sum of (dropped) probabilities: multiplies the value bias
    float t = 0.f;
    for (int jj = 0; jj < TOK_PER_CTA; ++jj) t += ps[tid][jj];
    red_add_f32(psum + b * GH + tid, t);
  }
  // thread owns columns {tid*2, tid*2+1} of each of the 768/512... use 3 column pairs: c = tid*2 + k*512? keep
  // coalescing simple: column pair index cp = k*256 + tid  (k = 0..1 covers 512 pairs > 384) -> use 384 pairs
  float acc[GH][2][2];
#pragma unroll
  for (int h = 0; h < GH; ++h) acc[h][0][0] = acc[h][0][1] = acc[h][1][0] = acc[h][1][1] = 0.f;
  const int n = min(TOK_PER_CTA, L - j0);
  for (int jj = 0; jj < n; ++jj) {
    const uint32_t* xr = reinterpret_cast<const uint32_t*>(x + (static_cast<size_t>(b) * L + j0 + jj) * GE);
    const float2 v0 = unpack_bf16(xr[tid]);
    const float2 v1 = (tid < 128) ? unpack_bf16(xr[256 + tid]) : make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < GH; ++h) {
      const float pw = ps[h][jj];
      acc[h][0][0] += pw * v0.x; acc[h][0][1] += pw * v0.y;
      acc[h][1][0] += pw * v1.x; acc[h][1][1] += pw * v1.y;
    }
  }
#pragma unroll
  for (int h = 0; h < GH; ++h) {
    float* mb = mvec + (static_cast<size_t>(b) * GH + h) * GE;
    red_add_f32(mb + tid * 2, acc[h][0][0]);
    red_add_f32(mb + tid * 2 + 1, acc[h][0][1]);
    if (tid < 128) {
      red_add_f32(mb + 512 + tid * 2, acc[h][1][0]);
      red_add_f32(mb + 512 + tid * 2 + 1, acc[h][1][1]);
    }
  }
}

// ---- G4: out[b, h*64+d] = Wvg[h*64+d,:] . m[b,h,:] + bvg * psum -> ctx row 0.  grid (H, B) ----
__global__ void __launch_bounds__(256)
global_out_kernel(const float* __restrict__ mvec, const float* __restrict__ psum, const float* __restrict__ Wvg,
                  const float* __restrict__ bvg, const uint8_t* __restrict__ mask, int L,
                  __nv_bfloat16* __restrict__ ctx) {
  const int h = blockIdx.x, b = blockIdx.y;
  if (mask[static_cast<size_t>(b) * L] != 2) return;   // no global token in this sequence
  __shared__ float ms[GE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* mb = mvec + (static_cast<size_t>(b) * GH + h) * GE;
  for (int e = tid; e < GE; e += 256) ms[e] = mb[e];
  __syncthreads();
  const float ps = psum[b * GH + h];
  for (int d = warp; d < GD; d += 8) {
    const float* wr = Wvg + static_cast<size_t>(h * GD + d) * GE;
    float acc = 0.f;
    for (int e = lane; e < GE; e += 32) acc += wr[e] * ms[e];
    acc = warp_sum(acc);
    if (lane == 0) ctx[static_cast<size_t>(b) * L * GE + h * GD + d] = __float2bfloat16(acc + bvg[h * GD + d] * ps);
  }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_global_attn_fwd(const rf_global_args* a, void* ctx, float* qg, float* u, float* p, float* mvec,
                                  float* psum, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && ctx && qg && u && p && mvec && psum, "rf_global_attn_fwd: null argument");
  RF_REQUIRE(a->H == GH && a->D == GD, "rf_global_attn_fwd: only H=12, D=64 supported (got %d, %d)", a->H, a->D);
  RF_REQUIRE(a->B > 0 && a->L > 0, "rf_global_attn_fwd: bad shape");
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  const int chunks = (a->L + TOK_PER_CTA - 1) / TOK_PER_CTA;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(global_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GH * GE * 4));
    attr_set = true;
  }
  global_qu_kernel<<<dim3(GH, a->B), 256, 0, stream>>>(x, a->mask012, a->Wqg, a->bqg, a->Wkg, a->L, qg, u, mvec, psum);
  int rc = check_launch("rf_global_attn_fwd/qu");
  if (rc) return rc;
  global_scores_kernel<<<dim3(chunks, a->B), 256, GH * GE * 4, stream>>>(x, a->mask012, u, a->L, p);
  rc = check_launch("rf_global_attn_fwd/scores");
  if (rc) return rc;
  global_softmax_kernel<<<a->B * GH, 256, 0, stream>>>(p, a->L);
  rc = check_launch("rf_global_attn_fwd/softmax");
  if (rc) return rc;
  const uint32_t thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  const float scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  global_mix_kernel<<<dim3(chunks, a->B), 256, 0, stream>>>(x, p, a->L, scale, thresh, a->drop_seed, mvec, psum);
  rc = check_launch("rf_global_attn_fwd/mix");
  if (rc) return rc;
  global_out_kernel<<<dim3(GH, a->B), 256, 0, stream>>>(mvec, psum, a->Wvg, a->bvg, a->mask012, a->L,
                                                       reinterpret_cast<__nv_bfloat16*>(ctx));
  return check_launch("rf_global_attn_fwd/out");
}
