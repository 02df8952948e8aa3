// Banded (sliding-window) attention forward with a global CLS key column, sm_100a.
//
// One CTA = one (batch, head, 128-query tile).  Keys j in [i0-W, i0+127+W] (NK = 128+2W rows)
// plus the sequence's first 16 rows (row 0 = the global CLS key) are TMA-loaded once
// (128B-swizzled, out-of-range rows zero-filled by TMA), then
//   S[128 x (NK+16)] = Q K^T          tcgen05.mma, fp32 accumulator in TMEM
//   softmax over the band + CLS column in fp32, one thread per query row (tcgen05.ld); the
//   band / padding / global masks are predicates on (row, column), never materialised
//   P (bf16) -> shared memory in the K-major UMMA layout (aliasing the dead Q/K tiles)
//   O[128 x 64] = P V                 tcgen05.mma (V is the MN-major B operand, no transpose)
//   O / rowsum -> bf16 context; log-sum-exp saved for the backward pass.
// Semantics: SURVEY.md §8a Spec A / HF:481-639.  Global keys are removed from the band and
// re-enter through the extra column (HF:523,558-568); padded query rows produce zeros (HF:578);
// the global query row (position 0 when mask012==2) is left to rf_global_attn_fwd (HF:963-1056).
#include <cuda_bf16.h>
#include <math.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

RF_DEFINE_NONCE_LOADER(attn_fwd)

namespace rf {

constexpr int ATT_THREADS = 128;
constexpr int HEAD_DIM = 64;

template <int W>
struct AttnFwdCfg {
  static constexpr int NK = 128 + 2 * W;       // band key rows
  static constexpr int NT = NK + 16;           // + global chunk
  static constexpr int PCH = (NT + 63) / 64;   // 64-key P chunks
  static constexpr uint32_t Q_BYTES = 128 * 128;
  static constexpr uint32_t KV_BYTES = NT * 128;
  static constexpr uint32_t P_BYTES = PCH * 16384;
  static constexpr uint32_t REGION_A = (Q_BYTES + KV_BYTES > P_BYTES) ? (Q_BYTES + KV_BYTES) : P_BYTES;
  static constexpr uint32_t OFF_V = REGION_A;
  static constexpr uint32_t OFF_FLAG = OFF_V + KV_BYTES;
  static constexpr uint32_t OFF_BAR = OFF_FLAG + ((NT + 15) / 16) * 16;
  static constexpr uint32_t TOTAL = OFF_BAR + 64 + 1024;
  static constexpr uint32_t TMEM_COLS = (NT <= 256) ? 256 : 512;
  static_assert(NT <= 512, "window too large for the single-shot kernel");
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

struct AttnFwdParams {
  const uint8_t* mask012;
  __nv_bfloat16* ctx;
  float* lse;
  int B, L, H;
  // window segment (windows wider than 2*W+1 keys are covered by several launches whose outputs are merged
  // through their log-sum-exps): keys are shifted by `shift` rows, the top `hi_cut` offsets of the band are
  // cut, the CLS column is only part of segment 0, and an empty row reports lse = -inf instead of 0
  int shift, hi_cut, use_cls, lse_neg_inf;
  float drop_scale;
  uint32_t drop_thresh;
  uint64_t drop_seed;
};

template <int W>
__global__ void __launch_bounds__(ATT_THREADS)
band_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
                     const AttnFwdParams p) {
  using C = AttnFwdCfg<W>;
  constexpr int NK = C::NK, NT = C::NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + C::Q_BYTES;
  uint8_t* sP = smem;  // aliases Q/K once S has been computed
  uint8_t* sV = smem + C::OFF_V;
  uint8_t* kflag = smem + C::OFF_FLAG;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int tiles_per_seq = (p.L + 127) / 128;
  const int tile = blockIdx.x % tiles_per_seq;
  const int h = (blockIdx.x / tiles_per_seq) % p.H;
  const int b = blockIdx.x / (tiles_per_seq * p.H);
  const int i0 = tile * 128;
  const int E = p.H * HEAD_DIM;
  const uint8_t* mrow = p.mask012 + static_cast<size_t>(b) * p.L;

  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
    // the operand loads are issued first: the TMEM allocation and the key-flag set-up below (global loads
    // of the mask) then run under their latency
    mbar_arrive_expect_tx(bar_load, C::Q_BYTES + 2 * C::KV_BYTES);
#pragma unroll
    for (int c = 0; c < 2; ++c) tma_load_3d(sQ + c * 8192, &tm64, bar_load, h * HEAD_DIM, i0 + c * 64, b);
#pragma unroll
    for (int c = 0; c < NK / 64; ++c) {
      tma_load_3d(sK + c * 8192, &tm64, bar_load, E + h * HEAD_DIM, i0 - W + p.shift + c * 64, b);
      tma_load_3d(sV + c * 8192, &tm64, bar_load, 2 * E + h * HEAD_DIM, i0 - W + p.shift + c * 64, b);
    }
    tma_load_3d(sK + NK * 128, &tm16, bar_load, E + h * HEAD_DIM, 0, b);
    tma_load_3d(sV + NK * 128, &tm16, bar_load, 2 * E + h * HEAD_DIM, 0, b);
  }
  // the row's mask bytes are loaded now, next to the TMA loads (first used after the S MMA)
  const uint8_t m_row = mrow[(i0 + tid) < p.L ? (i0 + tid) : 0], m_cls = mrow[0];
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  for (int c = tid; c < NT; c += ATT_THREADS) {
    uint8_t f = 0;
    if (c < NK) {
      const int j = i0 - W + p.shift + c;
      f = (j >= 0 && j < p.L && mrow[j] == 1) ? 1 : 0;
    } else if (c == NK) {
      f = (p.use_cls && mrow[0] == 2) ? 1 : 0;
    }
    kflag[c] = f;
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
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();

  // ---- softmax: thread r owns query row i0 + r (TMEM lane r) ----
  const int r = tid;
  const int i = i0 + r;
  const bool row_valid = (i < p.L) && (m_row != 0);
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  constexpr int WIN_CH = 2 * W / 32 + 1;   // 32-column chunks that can hold this warp's band
  const float LOG2E = 1.4426950408889634f;

  float m = -INFINITY;
  {
#pragma unroll 1
    for (int cc = warp; cc < warp + WIN_CH; ++cc) {
      uint32_t v[32];
      tmem_ld32(lane_base + cc * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = cc * 32 + j;
        const bool ok = kflag[c] && (c >= r) && (c <= r + band_hi);
        m = ok ? fmaxf(m, __uint_as_float(v[j])) : m;
      }
    }
  }
  uint32_t g16[16];
  tmem_ld16(lane_base + NK, g16);
  tmem_ld_wait();
  const float sg = __uint_as_float(g16[0]);
  const bool g_ok = kflag[NK] != 0;
  if (g_ok) m = fmaxf(m, sg);
  if (!row_valid || m == -INFINITY) m = 0.0f;   // fully masked row: every p below is forced to 0
  const float m2 = m * LOG2E;

  float l = 0.0f;
#pragma unroll 1
  for (int cc = 0; cc < NK / 32; ++cc) {
    uint4 out[4];
    if (cc >= warp && cc < warp + WIN_CH) {   // warp-uniform
      uint32_t v[32];
      tmem_ld32(lane_base + cc * 32, v);
      tmem_ld_wait();
      float pr[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = cc * 32 + j;
        const bool ok = row_valid && kflag[c] && (c >= r) && (c <= r + band_hi);
        const float e = ok ? exp2f(__uint_as_float(v[j]) * LOG2E - m2) : 0.0f;
        l += e;
        pr[j] = e;
      }
      if (p.drop_thresh != 0) {
        // keep-mask keyed on (row, 8-column group of the tile): one Philox call per 8 probabilities
        const uint64_t rowid = (static_cast<uint64_t>(b) * p.H + h) * p.L + i;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c0 = cc * 32 + u * 8;
          if (c0 + 7 >= r && c0 <= r + 2 * W) {
            const uint32_t keep = dropout_keep8(p.drop_seed, rowid * (NT / 8) + (c0 >> 3), p.drop_thresh);
#pragma unroll
            for (int e = 0; e < 8; ++e) pr[u * 8 + e] = ((keep >> e) & 1u) ? pr[u * 8 + e] * p.drop_scale : 0.0f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        out[u].x = pack_bf16(pr[u * 8 + 0], pr[u * 8 + 1]);
        out[u].y = pack_bf16(pr[u * 8 + 2], pr[u * 8 + 3]);
        out[u].z = pack_bf16(pr[u * 8 + 4], pr[u * 8 + 5]);
        out[u].w = pack_bf16(pr[u * 8 + 6], pr[u * 8 + 7]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) out[u] = make_uint4(0, 0, 0, 0);
    }
    uint8_t* prow = sP + (cc >> 1) * 16384 + r * 128;
    const int ubase = (cc & 1) * 4;
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(prow + (((ubase + u) ^ (r & 7)) << 4)) = out[u];
  }
  {
    float pg = (row_valid && g_ok) ? exp2f(sg * LOG2E - m2) : 0.0f;
    l += pg;
    if (p.drop_thresh != 0) {
      const uint64_t rowid = (static_cast<uint64_t>(b) * p.H + h) * p.L + i;
      const uint32_t keep = dropout_keep8(p.drop_seed, rowid * (NT / 8) + (NK >> 3), p.drop_thresh);
      pg = (keep & 1u) ? pg * p.drop_scale : 0.0f;
    }
    uint8_t* prow = sP + (NK / 64) * 16384 + r * 128;
    *reinterpret_cast<uint4*>(prow + ((0 ^ (r & 7)) << 4)) = make_uint4(pack_bf16(pg, 0.0f), 0, 0, 0);
    *reinterpret_cast<uint4*>(prow + ((1 ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
  }

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
  __syncwarp();
  mbar_wait(bar_mma, 1);
  tc_fence_after();

  const float inv_l = l > 0.0f ? 1.0f / l : 0.0f;
  const bool is_global_row = (i == 0) && (m_cls == 2);
  const bool do_store = (i < p.L) && !is_global_row;
  __nv_bfloat16* orow = p.ctx + (static_cast<size_t>(b) * p.L + (do_store ? i : 0)) * E + h * HEAD_DIM;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld32(lane_base + half * 32, v);   // warp-collective: executed by every lane
    tmem_ld_wait();
    if (do_store) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(v[j]) * inv_l, __uint_as_float(v[j + 1]) * inv_l);
        o.y = pack_bf16(__uint_as_float(v[j + 2]) * inv_l, __uint_as_float(v[j + 3]) * inv_l);
        o.z = pack_bf16(__uint_as_float(v[j + 4]) * inv_l, __uint_as_float(v[j + 5]) * inv_l);
        o.w = pack_bf16(__uint_as_float(v[j + 6]) * inv_l, __uint_as_float(v[j + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + half * 32 + j) = o;
      }
    }
  }
  if (i < p.L)
    p.lse[(static_cast<size_t>(b) * p.H + h) * p.L + i] = (l > 0.0f) ? (m + logf(l)) : (p.lse_neg_inf ? -INFINITY : 0.0f);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, C::TMEM_COLS);
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

// Segments of a band of half-width w for the W = 32 kernels (65 key offsets each): offsets [-w + 65k, -w + 65k + 64]
// clipped to [-w, w]; the kernel's own band is centred, so segment k is run with keys shifted by its centre.
static int attn_segments(int w, AttnSegment* seg) {
  const int n = (2 * w + 1 + 64) / 65;
  for (int k = 0; k < n; ++k) {
    const int lo = -w + 65 * k, hi = lo + 64;
    seg[k].shift = lo + 32;
    seg[k].hi_cut = hi > w ? hi - w : 0;
    seg[k].use_cls = k == 0;
  }
  return n;
}

static int launch_attn_fwd32(const rf_attn_args* a, void* ctx, float* lse, const AttnSegment& sg, int lse_neg_inf,
                             uint64_t seed, cudaStream_t stream) {
  constexpr int W = 32;
  using C = AttnFwdCfg<W>;
  auto kern = band_attn_fwd_kernel<W>;
  static bool attr_set = false;
  if (!attr_set) {
    RF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    attr_set = true;
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
  p.drop_seed = seed;
  const int tiles = (a->L + 127) / 128;
  kern<<<a->B * a->H * tiles, ATT_THREADS, C::TOTAL, stream>>>(*tm64, *tm16, p);
  return check_launch("rf_band_attn_fwd");
}

}  // namespace rf

extern "C" long long rf_band_attn_ws_bytes(int B, int L, int H, int w) {
  if (w <= 32) return 0;
  const long long T = static_cast<long long>(B) * L, E = static_cast<long long>(H) * rf::HEAD_DIM;
  // forward: fp32 accumulator [T,E] + bf16 segment output [T,E] + 2 x lse [B,H,L];  backward: fp32 dQ scratch [T,E]
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
  AttnSegment seg[16];
  const int nseg = attn_segments(a->w, seg);
  if (nseg == 1) return launch_attn_fwd32(a, ctx, lse, seg[0], 0, a->drop_seed, stream);
  RF_REQUIRE(a->ws != nullptr, "rf_band_attn_fwd: windows wider than 64 need a workspace (rf_band_attn_ws_bytes)");
  const size_t T = static_cast<size_t>(a->B) * a->L, E = static_cast<size_t>(a->H) * HEAD_DIM;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->ws);
  float* acc = reinterpret_cast<float*>(ws);
  __nv_bfloat16* part = reinterpret_cast<__nv_bfloat16*>(ws + T * E * 4);
  float* lse_part = reinterpret_cast<float*>(ws + T * E * 6);
  float* lse_acc = lse_part + static_cast<size_t>(a->B) * a->H * a->L;
  const long long total = static_cast<long long>(T) * a->H * 8;
  for (int k = 0; k < nseg; ++k) {
    int rc = launch_attn_fwd32(a, part, lse_part, seg[k], 1, a->drop_seed + 0x9E3779B97F4A7C15ull * k, stream);
    if (rc) return rc;
    attn_merge_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
        acc, lse_acc, part, lse_part, a->mask012, reinterpret_cast<__nv_bfloat16*>(ctx), lse, a->B, a->L, a->H, k == 0,
        k == nseg - 1);
    rc = check_launch("rf_band_attn_fwd/merge");
    if (rc) return rc;
  }
  return RF_OK;
}
