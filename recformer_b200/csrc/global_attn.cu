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
constexpr int TOK = 32;   // tokens per CTA in the token-parallel kernels (4 per warp)

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void load_row24(const __nv_bfloat16* __restrict__ row, int lane, float (&xv)[24]) {
  const uint4* xr = reinterpret_cast<const uint4*>(row);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const uint4 raw = xr[k * 32 + lane];
    const float2 a0 = unpack_bf16(raw.x), a1 = unpack_bf16(raw.y), a2 = unpack_bf16(raw.z), a3 = unpack_bf16(raw.w);
    xv[k * 8 + 0] = a0.x; xv[k * 8 + 1] = a0.y; xv[k * 8 + 2] = a1.x; xv[k * 8 + 3] = a1.y;
    xv[k * 8 + 4] = a2.x; xv[k * 8 + 5] = a2.y; xv[k * 8 + 6] = a3.x; xv[k * 8 + 7] = a3.y;
  }
}

// dot of a weight row (fp32, global) with a vector in shared memory; one warp, float4 loads
__device__ __forceinline__ float warp_dot768(const float* __restrict__ w, const float* xs, int lane) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w) + k * 32 + lane);
    const float4 b = *reinterpret_cast<const float4*>(xs + (k * 32 + lane) * 4);
    acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  return warp_sum(acc);
}

// ---- G1: q_g and u_h; also zeroes the m / psum accumulators.  grid (H, B), 256 threads ----
__global__ void __launch_bounds__(256)
global_qu_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ Wqg, const float* __restrict__ bqg,
                 const float* __restrict__ Wkg, int L, float* __restrict__ qg, float* __restrict__ u,
                 float* __restrict__ mvec, float* __restrict__ psum) {
  const int h = blockIdx.x, b = blockIdx.y;
  __shared__ __align__(16) float xs[GE];
  __shared__ float qs[GD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* xc = x + static_cast<size_t>(b) * L * GE;   // row 0 of the sequence
  for (int e = tid; e < GE; e += 256) {
    xs[e] = __bfloat162float(xc[e]);
    mvec[(static_cast<size_t>(b) * GH + h) * GE + e] = 0.f;
  }
  if (tid == 0) psum[b * GH + h] = 0.f;
  __syncthreads();
#pragma unroll 2
  for (int d = warp; d < GD; d += 8) {
    const float acc = warp_dot768(Wqg + static_cast<size_t>(h * GD + d) * GE, xs, lane);
    if (lane == 0) {
      const float q = (acc + bqg[h * GD + d]) * 0.125f;
      qs[d] = q;
      qg[static_cast<size_t>(b) * GE + h * GD + d] = q;
    }
  }
  __syncthreads();
  for (int e = tid; e < GE; e += 256) {
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < GD; ++d) acc += __ldg(Wkg + static_cast<size_t>(h * GD + d) * GE + e) * qs[d];
    u[(static_cast<size_t>(b) * GH + h) * GE + e] = acc;
  }
}

// ---- G2 / GB2: per-token dot products against 12 per-sequence vectors.  grid (L/32, B) ----
//   MODE 0: s[b,h,j] = u_h . x_j                       (-inf for padded keys)
//   MODE 1: dp[b,h,j] = keep_j*scale*(dm_h . x_j + dpsum_h)   (0 for padded keys)
template <int MODE>
__global__ void __launch_bounds__(256)
global_dots_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ mask,
                   const float* __restrict__ vecs, const float* __restrict__ dpsum, int L, float drop_scale,
                   uint32_t drop_thresh, uint64_t drop_seed, float* __restrict__ out) {
  const int b = blockIdx.y, j0 = blockIdx.x * TOK;
  extern __shared__ __align__(16) float us[];   // [GH][GE]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* ub = reinterpret_cast<const float4*>(vecs + static_cast<size_t>(b) * GH * GE);
  for (int i = tid; i < GH * GE / 4; i += 256) reinterpret_cast<float4*>(us)[i] = __ldg(ub + i);
  __syncthreads();
#pragma unroll 1
  for (int pair = 0; pair < 2; ++pair) {
    const int ja = j0 + warp * 4 + pair * 2, jb = ja + 1;
    if (ja >= L) break;
    const bool has_b = jb < L;
    float xa[24], xb[24];
    load_row24(x + (static_cast<size_t>(b) * L + ja) * GE, lane, xa);
    load_row24(x + (static_cast<size_t>(b) * L + (has_b ? jb : ja)) * GE, lane, xb);
    float ra[GH], rb[GH];
#pragma unroll
    for (int h = 0; h < GH; ++h) {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* up = us + h * GE + (k * 32 + lane) * 8;
        const float4 u0 = *reinterpret_cast<const float4*>(up), u1 = *reinterpret_cast<const float4*>(up + 4);
        sa += xa[k * 8 + 0] * u0.x + xa[k * 8 + 1] * u0.y + xa[k * 8 + 2] * u0.z + xa[k * 8 + 3] * u0.w +
              xa[k * 8 + 4] * u1.x + xa[k * 8 + 5] * u1.y + xa[k * 8 + 6] * u1.z + xa[k * 8 + 7] * u1.w;
        sb += xb[k * 8 + 0] * u0.x + xb[k * 8 + 1] * u0.y + xb[k * 8 + 2] * u0.z + xb[k * 8 + 3] * u0.w +
              xb[k * 8 + 4] * u1.x + xb[k * 8 + 5] * u1.y + xb[k * 8 + 6] * u1.z + xb[k * 8 + 7] * u1.w;
      }
      ra[h] = sa; rb[h] = sb;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int h = 0; h < GH; ++h) {
        ra[h] += __shfl_xor_sync(0xffffffffu, ra[h], o);
        rb[h] += __shfl_xor_sync(0xffffffffu, rb[h], o);
      }
    }
    if (lane < GH) {
      const int h = lane;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int j = t == 0 ? ja : jb;
        if (j >= L) break;
        float v = 0.f;
#pragma unroll
        for (int hh = 0; hh < GH; ++hh) if (hh == h) v = (t == 0) ? ra[hh] : rb[hh];
        const bool valid = mask[static_cast<size_t>(b) * L + j] != 0;
        if (MODE == 0) {
          v = valid ? v : -INFINITY;
        } else {
          v = valid ? v + dpsum[b * GH + h] : 0.f;
          if (drop_thresh != 0) {
            const uint64_t idx = (static_cast<uint64_t>(b) * GH + h) * L + j;
            const uint32_t keep = dropout_keep8(drop_seed, idx >> 3, drop_thresh);
            v = ((keep >> (idx & 7)) & 1u) ? v * drop_scale : 0.f;
          }
        }
        out[(static_cast<size_t>(b) * GH + h) * L + j] = v;
      }
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

// coefficient tile loader shared by mix / dx: cp[h][jj] = dropout(p)[b,h,j0+jj] (and cs = ds)
__device__ __forceinline__ void load_coefs(const float* __restrict__ p, const float* __restrict__ ds, int b, int j0,
                                           int L, float drop_scale, uint32_t drop_thresh, uint64_t drop_seed,
                                           float (*cp)[TOK], float (*cs)[TOK]) {
  for (int i = threadIdx.x; i < GH * TOK; i += 256) {
    const int h = i / TOK, jj = i % TOK;
    const int j = j0 + jj;
    float v = 0.f, s = 0.f;
    if (j < L) {
      const uint64_t idx = (static_cast<uint64_t>(b) * GH + h) * L + j;
      v = p[idx];
      if (ds) s = ds[idx];
      if (drop_thresh != 0) {
        const uint32_t keep = dropout_keep8(drop_seed, idx >> 3, drop_thresh);
        v = ((keep >> (idx & 7)) & 1u) ? v * drop_scale : 0.f;
      }
    }
    cp[h][jj] = v;
    if (cs) cs[h][jj] = s;
  }
}

// ---- G3: m[b,h,:] += sum_{j in chunk} p'[b,h,j] x[b,j,:].  grid (L/32, B), 256 threads ----
// thread owns column pair `tid` (cols 2tid, 2tid+1) and, for tid < 128, pair 256+tid
__global__ void __launch_bounds__(256)
global_mix_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ p, int L, float drop_scale,
                  uint32_t drop_thresh, uint64_t drop_seed, float* __restrict__ mvec, float* __restrict__ psum) {
  const int b = blockIdx.y, j0 = blockIdx.x * TOK;
  __shared__ float ps[GH][TOK];
  const int tid = threadIdx.x;
  load_coefs(p, nullptr, b, j0, L, drop_scale, drop_thresh, drop_seed, ps, nullptr);
  __syncthreads();
  if (tid < GH) {   // sum of (dropped) probabilities: multiplies the value bias
    float t = 0.f;
    for (int jj = 0; jj < TOK; ++jj) t += ps[tid][jj];
    red_add_f32(psum + b * GH + tid, t);
  }
  const bool second = tid < 128;
  float acc[GH][4];
#pragma unroll
  for (int h = 0; h < GH; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
  const int n = min(TOK, L - j0);
  for (int jj = 0; jj < n; jj += 4) {
    uint32_t r0[4], r1[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = min(jj + t, n - 1);
      const uint32_t* xr = reinterpret_cast<const uint32_t*>(x + (static_cast<size_t>(b) * L + j0 + j) * GE);
      r0[t] = xr[tid];
      r1[t] = second ? xr[256 + tid] : 0u;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (jj + t < n) {
        const float2 v0 = unpack_bf16(r0[t]), v1 = unpack_bf16(r1[t]);
#pragma unroll
        for (int h = 0; h < GH; ++h) {
          const float pw = ps[h][jj + t];
          acc[h][0] += pw * v0.x; acc[h][1] += pw * v0.y; acc[h][2] += pw * v1.x; acc[h][3] += pw * v1.y;
        }
      }
    }
  }
#pragma unroll
  for (int h = 0; h < GH; ++h) {
    float* mb = mvec + (static_cast<size_t>(b) * GH + h) * GE;
    red_add_f32(mb + tid * 2, acc[h][0]);
    red_add_f32(mb + tid * 2 + 1, acc[h][1]);
    if (second) {
      red_add_f32(mb + 512 + tid * 2, acc[h][2]);
      red_add_f32(mb + 512 + tid * 2 + 1, acc[h][3]);
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
  __shared__ __align__(16) float ms[GE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* mb = mvec + (static_cast<size_t>(b) * GH + h) * GE;
  for (int e = tid; e < GE; e += 256) ms[e] = mb[e];
  __syncthreads();
  const float ps = psum[b * GH + h];
#pragma unroll 2
  for (int d = warp; d < GD; d += 8) {
    const float acc = warp_dot768(Wvg + static_cast<size_t>(h * GD + d) * GE, ms, lane);
    if (lane == 0) ctx[static_cast<size_t>(b) * L * GE + h * GD + d] = __float2bfloat16(acc + bvg[h * GD + d] * ps);
  }
}

// =============================== backward of the CLS row ===================================
// GB1: dm[b,h,:] = Wvg[h]^T dout_h, dpsum = bvg[h].dout_h, dbvg += dout_h psum; keeps an fp32 copy
//      of dout (zero for sequences without a global token); zeroes du and dxcls.  grid (H, B)
__global__ void __launch_bounds__(256)
global_bwd_dm_kernel(const __nv_bfloat16* __restrict__ dctx, const uint8_t* __restrict__ mask,
                     const float* __restrict__ Wvg, const float* __restrict__ bvg, const float* __restrict__ psum,
                     int L, float* __restrict__ dm, float* __restrict__ dpsum, float* __restrict__ du,
                     float* __restrict__ dxcls, float* __restrict__ doutf, float* dbvg) {
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  __shared__ float dout[GD];
  const bool on = mask[static_cast<size_t>(b) * L] == 2;
  if (tid < GD) {
    const float g = on ? __bfloat162float(dctx[static_cast<size_t>(b) * L * GE + h * GD + tid]) : 0.f;
    dout[tid] = g;
    doutf[static_cast<size_t>(b) * GE + h * GD + tid] = g;
  }
  for (int e = tid; e < GE; e += 256) du[(static_cast<size_t>(b) * GH + h) * GE + e] = 0.f;
  if (h == 0) for (int e = tid; e < GE; e += 256) dxcls[static_cast<size_t>(b) * GE + e] = 0.f;
  __syncthreads();
  for (int e = tid; e < GE; e += 256) {
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < GD; ++d) acc += __ldg(Wvg + static_cast<size_t>(h * GD + d) * GE + e) * dout[d];
    dm[(static_cast<size_t>(b) * GH + h) * GE + e] = acc;
  }
  if (tid < GD && on && dbvg) red_add_f32(dbvg + h * GD + tid, dout[tid] * psum[b * GH + h]);
  if (tid == 0) {
    float t = 0.f;
    for (int d = 0; d < GD; ++d) t += bvg[h * GD + d] * dout[d];
    dpsum[b * GH + h] = t;
  }
}

// GB3: ds_j = p_j (dp_j - sum_k p_k dp_k), in place over dp.  grid (B*H), 256 threads
__global__ void __launch_bounds__(256) global_bwd_ds_kernel(const float* __restrict__ p, float* __restrict__ dp, int L) {
  const float* pr = p + static_cast<size_t>(blockIdx.x) * L;
  float* dr = dp + static_cast<size_t>(blockIdx.x) * L;
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
  for (int j = tid; j < L; j += 256) dr[j] = pr[j] * (dr[j] - t);
}

// GB4: du[b,h,:] += sum_j ds_hj x_j ;  dx_j += sum_h (p'_hj dm_h + ds_hj u_h).  grid (L/32, B)
__global__ void __launch_bounds__(256)
global_bwd_dx_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ p, const float* __restrict__ ds,
                     const float* __restrict__ dm, const float* __restrict__ u, int L, float drop_scale,
                     uint32_t drop_thresh, uint64_t drop_seed, float* __restrict__ du, __nv_bfloat16* __restrict__ dx) {
  const int b = blockIdx.y, j0 = blockIdx.x * TOK;
  __shared__ float cp[GH][TOK];   // p'
  __shared__ float cs[GH][TOK];   // ds
  const int tid = threadIdx.x;
  load_coefs(p, ds, b, j0, L, drop_scale, drop_thresh, drop_seed, cp, cs);
  float dmr[GH][4], ur[GH][4], acc[GH][4];
  const bool second = tid < 128;
#pragma unroll
  for (int h = 0; h < GH; ++h) {
    const float* dmb = dm + (static_cast<size_t>(b) * GH + h) * GE;
    const float* ub = u + (static_cast<size_t>(b) * GH + h) * GE;
    const float2 d0 = *reinterpret_cast<const float2*>(dmb + tid * 2), u0 = *reinterpret_cast<const float2*>(ub + tid * 2);
    dmr[h][0] = d0.x; dmr[h][1] = d0.y; ur[h][0] = u0.x; ur[h][1] = u0.y;
    if (second) {
      const float2 d1 = *reinterpret_cast<const float2*>(dmb + 512 + tid * 2);
      const float2 u1 = *reinterpret_cast<const float2*>(ub + 512 + tid * 2);
      dmr[h][2] = d1.x; dmr[h][3] = d1.y; ur[h][2] = u1.x; ur[h][3] = u1.y;
    } else {
      dmr[h][2] = dmr[h][3] = ur[h][2] = ur[h][3] = 0.f;
    }
    acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
  }
  __syncthreads();
  const int n = min(TOK, L - j0);
  for (int jj = 0; jj < n; jj += 4) {
    uint32_t xr0[4], xr1[4], gr0[4], gr1[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = min(jj + t, n - 1);
      const size_t rowoff = (static_cast<size_t>(b) * L + j0 + j) * GE;
      const uint32_t* xr = reinterpret_cast<const uint32_t*>(x + rowoff);
      const uint32_t* dxr = reinterpret_cast<const uint32_t*>(dx + rowoff);
      xr0[t] = xr[tid]; gr0[t] = dxr[tid];
      xr1[t] = second ? xr[256 + tid] : 0u;
      gr1[t] = second ? dxr[256 + tid] : 0u;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (jj + t < n) {
        const float2 x0 = unpack_bf16(xr0[t]), x1 = unpack_bf16(xr1[t]);
        float2 g0 = unpack_bf16(gr0[t]), g1 = unpack_bf16(gr1[t]);
#pragma unroll
        for (int h = 0; h < GH; ++h) {
          const float a = cp[h][jj + t], s = cs[h][jj + t];
          g0.x += a * dmr[h][0] + s * ur[h][0]; g0.y += a * dmr[h][1] + s * ur[h][1];
          g1.x += a * dmr[h][2] + s * ur[h][2]; g1.y += a * dmr[h][3] + s * ur[h][3];
          acc[h][0] += s * x0.x; acc[h][1] += s * x0.y; acc[h][2] += s * x1.x; acc[h][3] += s * x1.y;
        }
        uint32_t* dxw = reinterpret_cast<uint32_t*>(dx + (static_cast<size_t>(b) * L + j0 + jj + t) * GE);
        dxw[tid] = pack_bf16(g0.x, g0.y);
        if (second) dxw[256 + tid] = pack_bf16(g1.x, g1.y);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < GH; ++h) {
    float* dub = du + (static_cast<size_t>(b) * GH + h) * GE;
    red_add_f32(dub + tid * 2, acc[h][0]);
    red_add_f32(dub + tid * 2 + 1, acc[h][1]);
    if (second) {
      red_add_f32(dub + 512 + tid * 2, acc[h][2]);
      red_add_f32(dub + 512 + tid * 2 + 1, acc[h][3]);
    }
  }
}

// GB5: dq = Wkg[h] du_h / 8 (fp32 copy kept), dbqg += dq, dxcls += Wqg[h]^T dq.  grid (H, B)
__global__ void __launch_bounds__(256)
global_bwd_q_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ Wqg, const float* __restrict__ Wkg,
                    const float* __restrict__ du, int L, float* __restrict__ dqf, float* dbqg,
                    float* __restrict__ dxcls) {
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool on = mask[static_cast<size_t>(b) * L] == 2;
  __shared__ __align__(16) float dus[GE];
  __shared__ float dq[GD];
  const float* dub = du + (static_cast<size_t>(b) * GH + h) * GE;
  for (int e = tid; e < GE; e += 256) dus[e] = dub[e];
  __syncthreads();
#pragma unroll 2
  for (int d = warp; d < GD; d += 8) {
    const float acc = warp_dot768(Wkg + static_cast<size_t>(h * GD + d) * GE, dus, lane);
    if (lane == 0) {
      const float g = on ? acc * 0.125f : 0.f;   // gradient w.r.t. (Wqg x + bqg)
      dq[d] = g;
      dqf[static_cast<size_t>(b) * GE + h * GD + d] = g;
      if (on && dbqg) red_add_f32(dbqg + h * GD + d, g);
    }
  }
  __syncthreads();
  if (!on) return;
  for (int e = tid; e < GE; e += 256) {
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < GD; ++d) acc += __ldg(Wqg + static_cast<size_t>(h * GD + d) * GE + e) * dq[d];
    red_add_f32(dxcls + static_cast<size_t>(b) * GE + e, acc);
  }
}

// GB6: dx[b,0,:] += dxcls[b,:].  grid (B), 256 threads
__global__ void __launch_bounds__(256)
global_bwd_cls_kernel(const float* __restrict__ dxcls, const uint8_t* __restrict__ mask, int L,
                      __nv_bfloat16* __restrict__ dx) {
  const int b = blockIdx.x;
  if (mask[static_cast<size_t>(b) * L] != 2) return;
  for (int e = threadIdx.x; e < GE; e += 256) {
    __nv_bfloat16* d = dx + static_cast<size_t>(b) * L * GE + e;
    *d = __float2bfloat16(__bfloat162float(*d) + dxcls[static_cast<size_t>(b) * GE + e]);
  }
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
    for (int bb = 0; bb < nb; ++bb) {
      const int b = b0 + bb;
      float v;
      if (z == 0) v = mvec[(static_cast<size_t>(b) * GH + h) * GE + e];
      else if (z == 1) v = du[(static_cast<size_t>(b) * GH + h) * GE + e];
      else v = __bfloat162float(x[static_cast<size_t>(b) * L * GE + e]);
#pragma unroll
      for (int d = 0; d < GD; ++d) acc[d] += as[bb][d] * v;
    }
  }
#pragma unroll
  for (int d = 0; d < GD; ++d) dW[static_cast<size_t>(h * GD + d) * GE + e] += acc[d];
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
  const int chunks = (a->L + TOK - 1) / TOK;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(global_dots_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GH * GE * 4));
    RF_CUDA(cudaFuncSetAttribute(global_dots_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GH * GE * 4));
    attr_set = true;
  }
  const uint32_t thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  const float scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  global_qu_kernel<<<dim3(GH, a->B), 256, 0, stream>>>(x, a->Wqg, a->bqg, a->Wkg, a->L, qg, u, mvec, psum);
  int rc = check_launch("rf_global_attn_fwd/qu");
  if (rc) return rc;
  global_dots_kernel<0><<<dim3(chunks, a->B), 256, GH * GE * 4, stream>>>(x, a->mask012, u, nullptr, a->L, scale,
                                                                        thresh, a->drop_seed, p);
  rc = check_launch("rf_global_attn_fwd/scores");
  if (rc) return rc;
  global_softmax_kernel<<<a->B * GH, 256, 0, stream>>>(p, a->L);
  rc = check_launch("rf_global_attn_fwd/softmax");
  if (rc) return rc;
  global_mix_kernel<<<dim3(chunks, a->B), 256, 0, stream>>>(x, p, a->L, scale, thresh, a->drop_seed, mvec, psum);
  rc = check_launch("rf_global_attn_fwd/mix");
  if (rc) return rc;
  global_out_kernel<<<dim3(GH, a->B), 256, 0, stream>>>(mvec, psum, a->Wvg, a->bvg, a->mask012, a->L,
                                                       reinterpret_cast<__nv_bfloat16*>(ctx));
  return check_launch("rf_global_attn_fwd/out");
}

extern "C" long long rf_global_attn_bwd_ws_bytes(int B, int L, int H) {
  const long long E = static_cast<long long>(H) * GD;
  return 4ll * (2ll * B * H * E + B * H + static_cast<long long>(B) * H * L + 3ll * B * E) + 256;
}

extern "C" int rf_global_attn_bwd(const rf_global_args* a, const void* dctx, const float* qg, const float* u,
                                  const float* p, const float* mvec, const float* psum, void* dx, float* dWqg,
                                  float* dbqg, float* dWkg, float* dWvg, float* dbvg, float* ws, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a && dctx && qg && u && p && mvec && psum && dx && ws, "rf_global_attn_bwd: null argument");
  RF_REQUIRE(a->H == GH && a->D == GD, "rf_global_attn_bwd: only H=12, D=64 supported (got %d, %d)", a->H, a->D);
  const int B = a->B, L = a->L;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  float* dm = ws;
  float* du = dm + static_cast<size_t>(B) * GH * GE;
  float* dpsum = du + static_cast<size_t>(B) * GH * GE;
  float* dp = dpsum + static_cast<size_t>(B) * GH;
  float* dxcls = dp + static_cast<size_t>(B) * GH * L;
  float* doutf = dxcls + static_cast<size_t>(B) * GE;
  float* dqf = doutf + static_cast<size_t>(B) * GE;
  const int chunks = (L + TOK - 1) / TOK;
  const uint32_t thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  const float scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(global_dots_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GH * GE * 4));
    RF_CUDA(cudaFuncSetAttribute(global_dots_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GH * GE * 4));
    attr_set = true;
  }
  global_bwd_dm_kernel<<<dim3(GH, B), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dctx), a->mask012,
                                                       a->Wvg, a->bvg, psum, L, dm, dpsum, du, dxcls, doutf, dbvg);
  int rc = check_launch("rf_global_attn_bwd/dm");
  if (rc) return rc;
  global_dots_kernel<1><<<dim3(chunks, B), 256, GH * GE * 4, stream>>>(x, a->mask012, dm, dpsum, L, scale, thresh,
                                                                     a->drop_seed, dp);
  rc = check_launch("rf_global_attn_bwd/dp");
  if (rc) return rc;
  global_bwd_ds_kernel<<<B * GH, 256, 0, stream>>>(p, dp, L);
  rc = check_launch("rf_global_attn_bwd/ds");
  if (rc) return rc;
  global_bwd_dx_kernel<<<dim3(chunks, B), 256, 0, stream>>>(x, p, dp, dm, u, L, scale, thresh, a->drop_seed, du,
                                                           reinterpret_cast<__nv_bfloat16*>(dx));
  rc = check_launch("rf_global_attn_bwd/dx");
  if (rc) return rc;
  global_bwd_q_kernel<<<dim3(GH, B), 256, 0, stream>>>(a->mask012, a->Wqg, a->Wkg, du, L, dqf, dbqg, dxcls);
  rc = check_launch("rf_global_attn_bwd/q");
  if (rc) return rc;
  global_bwd_cls_kernel<<<B, 256, 0, stream>>>(dxcls, a->mask012, L, reinterpret_cast<__nv_bfloat16*>(dx));
  rc = check_launch("rf_global_attn_bwd/cls");
  if (rc) return rc;
  global_wgrad_kernel<<<dim3(GH, 3, 3), 256, 0, stream>>>(x, a->mask012, B, L, doutf, mvec, qg, du, dqf, dWvg, dWkg,
                                                         dWqg);
  return check_launch("rf_global_attn_bwd/wgrad");
}
