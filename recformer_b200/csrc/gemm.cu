// Persistent warp-specialised bf16 GEMM on tcgen05 tensor cores for sm_100a.
//
//   C[M,N] = epilogue( A (*) B ),  fp32 accumulation in TMEM.
//
// One CTA per SM loops over 128 x BN output tiles (x split-K slices).  Roles:
//   warp 0      TMA producer: streams 128x64 A and BNx64 B slabs (128B swizzle) through a
//               STAGES-deep shared-memory ring, mbarrier complete_tx signalling;
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases ring
//               slots and publishes finished accumulators;
//   warps 2..5  epilogue: tcgen05.ld the accumulator (double-buffered in TMEM so the next
//               tile's MMAs overlap), apply bias / scale / GELU / GELU' / dropout / residual,
//               store bf16 or fp32 (plain, += or red.add for split-K).
// Operand layouts (template): K-major (row = M/N index, K contiguous: activations and
// nn.Linear weights) or MN-major (row = K index: used by dgrad for W and by wgrad for both
// operands, so no transposed copies of weights or activations are ever materialised).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rf_common.h"
#include "rf_ptx.cuh"

RF_DEFINE_NONCE_LOADER(gemm)

namespace rf {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 320;      // TMA warp + MMA warp + 8 epilogue warps
constexpr int EPI_WARPS = 8;

struct GemmParams {
  void* C;
  void* C2;
  const float* bias;
  const void* residual;     // bf16 or fp32 (residual_f32)
  int residual_f32;
  const __nv_bfloat16* aux;
  int M, N, K;
  int ldc, ldr, ldaux;
  int accumulate;
  int split_k;
  float scale;
  int scale_ncols;
  int debug_skip_epilogue;   // profiling aid (RF_DEBUG_GEMM_NOEPI=1): accumulators are drained but nothing is stored
  float drop_scale;       // 1/(1-p)
  uint32_t drop_thresh;   // p * 65536, 0 = no dropout
  uint64_t drop_seed;
  uint8_t* drop_mask;     // optional [M, N/8]: the dropout mask the epilogue draws, saved for rf_layernorm_bwd
  int xk_rows;            // XK: rows of A per batch (one 64-row block of B2 per batch)
  int* tile_counter;      // pair kernel: dynamic tile scheduler (zero between launches, see gemm_pair_kernel)
  // pair kernel, row activity (rf_set_row_activity): one flag per 256 rows of the token axis; 0 = padding only.
  const uint8_t* m_active;   // token axis = M (forward / dgrad): output tiles of inactive row tiles are skipped
  const uint8_t* k_active;   // token axis = K (wgrad): inactive 64-token k-blocks are skipped (their dY rows are zero)
};

template <int BN>
struct GemmSmem {
  static constexpr int STAGES = (BN == 256) ? 3 : 5;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TILE_BYTES = STAGES * STAGE_BYTES;
  static constexpr uint32_t BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr uint32_t BIAS_BYTES = 2 * BN * 4;                // per-accumulator-stage bias slab
  static constexpr uint32_t EPI_STAGE_BYTES = EPI_WARPS * 4096;     // per-warp transpose slab
  static constexpr uint32_t TOTAL = TILE_BYTES + EPI_STAGE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;  // + alignment slack
};

// GELU (erf form, HF ACT2FN["gelu"]) and its derivative from ONE exponential and one reciprocal:
// erf by Abramowitz & Stegun 7.1.25 (three coefficients, |error| <= 2.5e-5: the outputs are rounded to bf16, whose
// resolution is 4e-3 relative; the five-coefficient 7.1.26 used before bought nothing and cost two more FMAs),
//   erf(z) = 1 - (a1 t + a2 t^2 + a3 t^3) exp(-z^2),  t = 1/(1 + p z),  z = |x|/sqrt(2),
// and exp(-z^2) = exp(-x^2/2) is also the Gaussian of gelu'(x) = Phi(x) + x phi(x).  Two MUFU ops and ~14 FP32
// instructions per element: the epilogue is FP32-issue bound, every instruction counts.
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.47047f, z, 1.0f));
  float poly = fmaf(0.7478556f, t, -0.0958798f);
  poly = fmaf(poly, t, 0.3480242f);
  const float e = exp2f(-0.72134752044448170f * x * x);      // exp(-x^2/2)
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143268f, e, cdf);
}

// Epilogue of one 32-row x 32-column accumulator chunk held by a warp (thread = row, as read from
// TMEM).  The raw fp32 accumulators are first transposed through a 4 KB 128B-swizzled staging slab
// in shared memory so that every global access of the epilogue is coalesced: in the second phase
// lane l owns 8 consecutive columns ((l % 4) * 8) of row 8*s + l/4 for s = 0..3, i.e. each warp
// instruction touches 8 rows x 64..128 contiguous bytes instead of 32 rows x 16 bytes.
// Then: bias / scale / GELU / GELU' / dropout / residual, and the store.
//
// The body is STRAIGHT-LINE code: what the epilogue does is fixed by template flags (RES: 0 none, 1 bf16
// residual, 2 fp32 residual — also used for "accumulate into C"; DROP; SPLITK = red.add stores), rows past M
// are handled by clamped loads and predicated stores.  With run-time `if (residual) / if (dropout) /
// continue` the four row steps of a chunk were separate basic blocks and every shared / global load latency
// was exposed once per step; branch-free, the scheduler overlaps all four.
template <int EPI, bool OUT_F32, int RES, bool DROP, bool SPLITK>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], const GemmParams& p, const float* sb,
                                               uint8_t* stage, int lane, int lcol, int row_base, int col0) {
  const int cq = lane & 3;               // which 8-column group of the chunk
  const int c8 = col0 + cq * 8;          // first global column owned by this lane
  // ---- phase 0: issue the epilogue's global loads (residual / gelu' tile) for all four row steps ----
  uint4 aux_pf[4];
  float4 res_pf[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int rc = min(row_base + s * 8 + (lane >> 2), p.M - 1);
    if (EPI == RF_EPI_DGELU) aux_pf[s] = *reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(rc) * p.ldaux + c8);
    if (RES == 2) {
      const float* rs = reinterpret_cast<const float*>(p.residual) + static_cast<size_t>(rc) * p.ldr + c8;
      res_pf[s][0] = *reinterpret_cast<const float4*>(rs);
      res_pf[s][1] = *reinterpret_cast<const float4*>(rs + 4);
    } else if (RES == 1) {
      res_pf[s][0] = *reinterpret_cast<const float4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) +
                                                      static_cast<size_t>(rc) * p.ldr + c8);
    }
  }
  // ---- phase 1: thread (= row `lane`) writes its 32 fp32 values, 16B units XOR-swizzled by row ----
  {
    uint8_t* srow = stage + lane * 128;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      *reinterpret_cast<uint4*>(srow + ((u ^ (lane & 7)) << 4)) = make_uint4(r[u * 4], r[u * 4 + 1], r[u * 4 + 2], r[u * 4 + 3]);
  }
  __syncwarp();
  // ---- phase 2: coalesced layout (the bias slab is zero-filled when there is no bias) ----
  const float4 b0 = *reinterpret_cast<const float4*>(sb + lcol + cq * 8);
  const float4 b1 = *reinterpret_cast<const float4*>(sb + lcol + cq * 8 + 4);
  const float bias8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float sc = (col0 < p.scale_ncols) ? p.scale : 1.0f;   // scale_ncols is a multiple of 32
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int rl = s * 8 + (lane >> 2);
    const int row = row_base + rl;
    const bool ok = row < p.M;
    const uint8_t* srow = stage + rl * 128;
    const float4 x0 = *reinterpret_cast<const float4*>(srow + (((2 * cq) ^ (rl & 7)) << 4));
    const float4 x1 = *reinterpret_cast<const float4*>(srow + (((2 * cq + 1) ^ (rl & 7)) << 4));
    float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (EPI == RF_EPI_NONE) ? (v[e] + bias8[e]) * sc : v[e] + bias8[e];   // scale: plain epilogue only
    const size_t off = static_cast<size_t>(row) * p.ldc + c8;
    if (EPI == RF_EPI_GELU) {
      // C2 <- gelu(u) (operand of the next GEMM); C <- gelu'(u) (all the backward pass needs of u)
      float g[8], d[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) gelu_and_grad(v[e], g[e], d[e]);
      uint4 dv, gv;
      dv.x = pack_bf16(d[0], d[1]); dv.y = pack_bf16(d[2], d[3]); dv.z = pack_bf16(d[4], d[5]); dv.w = pack_bf16(d[6], d[7]);
      gv.x = pack_bf16(g[0], g[1]); gv.y = pack_bf16(g[2], g[3]); gv.z = pack_bf16(g[4], g[5]); gv.w = pack_bf16(g[6], g[7]);
      if (ok) {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off) = dv;
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C2) + off) = gv;
      }
    } else {
      if (EPI == RF_EPI_DGELU) {   // aux holds gelu'(u) saved by the forward epilogue
        const uint4 a = aux_pf[s];
        const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
        v[0] *= a0.x; v[1] *= a0.y; v[2] *= a1.x; v[3] *= a1.y; v[4] *= a2.x; v[5] *= a2.y; v[6] *= a3.x; v[7] *= a3.y;
      }
      if (DROP) {
        const uint64_t grp = (static_cast<uint64_t>(row) * p.N + c8) >> 3;
        const uint32_t keep = dropout_keep8(p.drop_seed, grp, p.drop_thresh);
        if (p.drop_mask != nullptr && ok) p.drop_mask[grp] = static_cast<uint8_t>(keep);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = ((keep >> e) & 1u) ? v[e] * p.drop_scale : 0.0f;
      }
      if (RES == 2) {
        const float4 r0 = res_pf[s][0], r1 = res_pf[s][1];
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      } else if (RES == 1) {
        const float4 raw = res_pf[s][0];
        const float2 a0 = unpack_bf16(__float_as_uint(raw.x)), a1 = unpack_bf16(__float_as_uint(raw.y));
        const float2 a2 = unpack_bf16(__float_as_uint(raw.z)), a3 = unpack_bf16(__float_as_uint(raw.w));
        v[0] += a0.x; v[1] += a0.y; v[2] += a1.x; v[3] += a1.y; v[4] += a2.x; v[5] += a2.y; v[6] += a3.x; v[7] += a3.y;
      }
      if (OUT_F32) {
        float* cf = reinterpret_cast<float*>(p.C) + off;
        if (SPLITK) {
          if (ok) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cf), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                         : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cf + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]),
                         "f"(v[7])
                         : "memory");
          }
        } else if (ok) {
          *reinterpret_cast<float4*>(cf) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(cf + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      } else {
        uint4 o;
        o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
        if (ok) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off) = o;
      }
    }
  }
  __syncwarp();   // the staging slab is rewritten by the next chunk
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool OUT_F32, int RES, bool DROP, bool SPLITK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_stage = smem + S::TILE_BYTES;    // [EPI_WARPS][4096] epilogue transpose slabs
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::TILE_BYTES + S::EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + S::TILE_BYTES + S::EPI_STAGE_BYTES + S::BAR_BYTES);   // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int k_blocks_total = (p.K + BK - 1) / BK;
  const int k_per_split = (k_blocks_total + p.split_k - 1) / p.split_k;
  const int total_tiles = m_tiles * n_tiles * p.split_k;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        const int rest = tile / n_tiles;
        const int mt = rest % m_tiles;
        const int sp = rest / m_tiles;
        const int m0 = mt * BM, n0 = nt * BN;
        const int kb0 = sp * k_per_split;
        const int kb1 = min(kb0 + k_per_split, k_blocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);            // box (64 k, 128 rows)
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)                                   // box (64 m, 64 k-rows)
              tma_load_2d(sa + c * 8192, &tmA, &full_bar[stage], m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);            // box (64 k, BN rows)
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * 8192, &tmB, &full_bar[stage], n0 + c * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
    const bool elected = elect_one();
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t adesc0 = A_MN ? umma_smem_desc(smem_base, 8192, 1024) : umma_smem_desc(smem_base, 16, 1024);
    const uint64_t bdesc0 = B_MN ? umma_smem_desc(smem_base + S::A_BYTES, 8192, 1024)
                                 : umma_smem_desc(smem_base + S::A_BYTES, 16, 1024);
    constexpr uint32_t A_KSTEP = (A_MN ? 2048u : 32u) >> 4, B_KSTEP = (B_MN ? 2048u : 32u) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int sp = (tile / n_tiles) / m_tiles;
      const int kb0 = sp * k_per_split;
      const int kb1 = min(kb0 + k_per_split, k_blocks_total);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t soff = static_cast<uint32_t>(stage) * (S::STAGE_BYTES >> 4);
        const uint64_t ad = adesc0 + soff, bd = bdesc0 + soff;
        if (elected) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, ad + k * A_KSTEP, bd + k * B_KSTEP, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);                 // slot reusable once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (kb1 <= kb0 && elected) umma_commit(&tfull_bar[acc]);  // empty K slice (never for sane shapes)
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue warps =================
    // warp (2..9): TMEM lane quadrant = warp % 4 (hardware rule), column half = (warp - 2) / 4
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;          // 0..255 among the epilogue threads
    constexpr int CH = BN / 2 / 32;             // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % n_tiles;
      const int mt = (tile / n_tiles) % m_tiles;
      const int m0 = mt * BM, n0 = nt * BN;
      const int row_base = m0 + quad * 32;
      uint8_t* my_stage = s_stage + (warp - 2) * 4096;
      // stage this tile's bias slab in shared memory (global-load latency off the critical path)
      float* sb = s_bias + acc * BN;
      if (etid < BN) sb[etid] = (p.bias != nullptr && n0 + etid < p.N) ? __ldg(p.bias + n0 + etid) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * (BN / 2);
      uint32_t rbuf[2][32];
      tmem_ld32(tbase, rbuf[0]);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        tmem_ld_wait();
        if (c + 1 < CH) {
          tmem_ld32(tbase + (c + 1) * 32, rbuf[(c + 1) & 1]);   // prefetch the next chunk
        } else {
          // accumulator fully read: hand the TMEM stage back to the MMA warp before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        const int lcol = half * (BN / 2) + c * 32;
        const int col0 = n0 + lcol;
        if (col0 >= p.N) continue;  // warp-uniform
        epilogue_chunk<EPI, OUT_F32, RES, DROP, SPLITK>(rbuf[c & 1], p, sb, my_stage, lane, lcol, row_base, col0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}



static void fill_params(const rf_gemm_args* a, GemmParams& p) {
  p.C = a->C; p.C2 = a->C2; p.bias = a->bias;
  p.residual = a->residual;
  p.residual_f32 = a->residual_f32;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux);
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.ldc = a->ldc; p.ldr = a->ldr; p.ldaux = a->ldaux;
  p.accumulate = a->accumulate;
  if (a->accumulate && a->out_f32 && a->residual == nullptr) {   // C += v  ==  v + (fp32 residual = C), read up front
    p.residual = a->C; p.residual_f32 = 1; p.ldr = a->ldc;
  }
  {
    // every K slice must be non-empty: recompute the slice count from the per-slice block count
    const int kb = (a->K + BK - 1) / BK;
    int sk = a->split_k > 1 ? a->split_k : 1;
    if (sk > kb) sk = kb;
    const int per = (kb + sk - 1) / sk;
    p.split_k = (kb + per - 1) / per;
  }
  p.scale = a->scale; p.scale_ncols = a->scale_ncols;
  p.drop_thresh = a->drop_p > 0.f ? static_cast<uint32_t>(a->drop_p * 65536.0f) : 0u;
  p.drop_scale = a->drop_p > 0.f ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  p.drop_seed = a->drop_seed;
  p.drop_mask = (a->drop_p > 0.f && a->N % 8 == 0) ? reinterpret_cast<uint8_t*>(a->drop_mask) : nullptr;
  static const int noepi = getenv("RF_DEBUG_GEMM_NOEPI") ? atoi(getenv("RF_DEBUG_GEMM_NOEPI")) : 0;
  p.debug_skip_epilogue = noepi;
}

// ==============================================================================================
// CTA-pair kernel: a 2-CTA cluster computes one 256 x 256 tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and its own 128-row half of B (32 KB per 64-wide K slab
// instead of 48 KB), so the per-SM L2->SM ingest that bounds the single-CTA kernel drops by a third
// and 6 ring stages fit.  The leader CTA's MMA thread issues for both; tcgen05.commit multicasts
// the "slot free" / "accumulator ready" arrivals to both CTAs; both epilogues report "accumulator
// drained" to the leader.
// ==============================================================================================
constexpr int P_BN = 256;        // pair tile N
constexpr int P_STAGES = 5;
constexpr uint32_t P_A_BYTES = 128 * BK * 2, P_B_BYTES = 128 * BK * 2, P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr uint32_t P_TILE_BYTES = P_STAGES * P_STAGE_BYTES;
constexpr int P_SCHED = 4;       // depth of the tile-id ring of the dynamic scheduler
constexpr uint32_t P_BAR_BYTES = (2 * P_STAGES + 4) * 8 + 16 + 2 * P_SCHED * 8 + P_SCHED * 4;
constexpr uint32_t P_EPI_STAGE_BYTES = EPI_WARPS * 4096;
constexpr uint32_t P_SMEM = P_TILE_BYTES + P_EPI_STAGE_BYTES + P_BAR_BYTES + 2 * P_BN * 4 + 1024;

// Tile scheduling is DYNAMIC: the leader's producer thread draws the next tile index from a global counter
// (atomicAdd) and publishes it through a 4-deep ring of tile ids in the shared memory of BOTH CTAs (a
// st.shared::cluster + release/acquire mbarrier hand-over); every role reads its tiles from that ring.  With the
// former static schedule (tile = pair, pair + npairs, ...) a pair whose SMs were still occupied when the grid was
// launched — by a kernel of another stream: the side-stream global-attention row, NCCL's all-reduce kernels in
// data-parallel training — started late and ran its whole share AFTER the others had finished, doubling the GEMM's
// duration; now the resident pairs drain the tile list and a late pair finds it (nearly) empty.  The counter is one
// of 256 zero-initialised device words picked round-robin by the host; the pair that draws the last sentinel
// (value total_tiles + npairs - 1: every pair draws exactly one) resets it to zero for its next use.
__device__ int g_gemm_tile_counters[256];

// wgrad with row activity: k-block kb (64 tokens) belongs to the 256-token tile kb >> 2
__device__ __forceinline__ bool kb_active(const uint8_t* k_active, int kb) {
  return k_active == nullptr || k_active[kb >> 2] != 0;
}
__device__ __forceinline__ int count_active_kb(const uint8_t* k_active, int kb0, int kb1) {
  if (k_active == nullptr) return kb1 - kb0;
  int n = 0;
  for (int kb = kb0; kb < kb1; ++kb) n += k_active[kb >> 2] != 0 ? 1 : 0;
  return n;
}

// XK: one extra 64-deep k-block per output tile whose operands come from a second pair of tensors, the A
// side [M, 64] K-major and the B side batched ([M / xk_rows] blocks of [64, N], N contiguous):
//   C = A B + A2[rows] B2[batch(rows)]  -- a per-sequence low-rank update riding on the dense GEMM
// (the global CLS row's token gradients in the QKV dgrad).
template <bool A_MN, bool B_MN, int EPI, bool OUT_F32, int RES, bool DROP, bool SPLITK, bool XK = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* s_stage = smem + P_TILE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P_TILE_BYTES + P_EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tfull_bar = empty_bar + P_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* sched_full = reinterpret_cast<uint64_t*>(tmem_slot + 4);
  uint64_t* sched_empty = sched_full + P_SCHED;      // the LEADER's copy is the live one
  volatile int* s_tile = reinterpret_cast<volatile int*>(sched_empty + P_SCHED);
  float* s_bias = reinterpret_cast<float*>(smem + P_TILE_BYTES + P_EPI_STAGE_BYTES + P_BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = gridDim.x >> 1;

  const int m_tiles = (p.M + 255) / 256;
  const int n_tiles = (p.N + P_BN - 1) / P_BN;
  const int k_blocks_total = (p.K + BK - 1) / BK;
  const int k_per_split = (k_blocks_total + p.split_k - 1) / p.split_k;
  const int total_tiles = m_tiles * n_tiles * p.split_k;

  // the scheduler thread draws its first tile before anything else: the atomic's round trip to L2 (~1 us) then
  // overlaps the barrier / TMEM set-up and the cluster sync below instead of delaying the first TMA load
  int first_draw = 0;
  if (threadIdx.x == 0 && leader && p.tile_counter != nullptr) first_draw = atomicAdd(p.tile_counter, 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);     // leader's copy is the live one: expect_tx covers both CTAs' slabs
      mbar_init(&empty_bar[s], 1);    // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);                // multicast tcgen05.commit
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);   // leader's copy: epilogue warps of both CTAs
    }
    for (int s = 0; s < P_SCHED; ++s) {
      mbar_init(&sched_full[s], 1);                    // the scheduler's publish (one arrive per CTA)
      mbar_init(&sched_empty[s], 2 * EPI_WARPS + 2);   // readers: 16 epilogue warps, the MMA warp, the peer's producer
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  // ring readers: wait for slot `rs`, read the tile id, report the read to the leader's sched_empty barrier
  int rs = 0;
  uint32_t rph = 0;
  auto read_tile = [&]() -> int {     // executed by whole warps (or by the peer's single producer thread)
    mbar_wait_cluster(&sched_full[rs], rph);
    return s_tile[rs];
  };
  auto release_tile = [&]() {         // one thread per reader, after the tile id has been consumed
    mbar_arrive_cluster(mapa_u32(smem_u32(&sched_empty[rs]), 0));
  };
  auto advance_ring = [&]() {
    if (++rs == P_SCHED) { rs = 0; rph ^= 1; }
  };
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      // leader: draw the next tile and publish it to both CTAs; peer: read it from the ring
      int static_next = static_cast<int>(blockIdx.x >> 1);
      bool have_first = true;
      auto next_tile = [&]() -> int {
        if (!leader) {
          const int t = read_tile();
          if (t >= 0) release_tile();      // (always true: orders the read before the arrive)
          advance_ring();
          return t;
        }
        mbar_wait(&sched_empty[rs], rph ^ 1);
        int t;
        if (p.tile_counter != nullptr) {
          t = have_first ? first_draw : atomicAdd(p.tile_counter, 1);
          have_first = false;
          if (t == total_tiles + npairs - 1) atomicExch(p.tile_counter, 0);   // last sentinel of this launch
        } else {                        // RF_GEMM_STATIC_SCHEDULE=1 (A/B aid): the former static round-robin
          t = static_next;
          static_next += npairs;
        }
        s_tile[rs] = t;
        st_shared_cluster_u32(mapa_u32(smem_u32(const_cast<int*>(&s_tile[rs])), 1), static_cast<uint32_t>(t));
        mbar_arrive(&sched_full[rs]);
        mbar_arrive_cluster_release(mapa_u32(smem_u32(&sched_full[rs]), 1));
        advance_ring();
        return t;
      };
      int tile = next_tile();
      while (tile < total_tiles) {
        const int tile_next = next_tile();      // drawn one tile ahead: its latency hides under this tile's loads
        const int nt = tile % n_tiles;
        const int rest = tile / n_tiles;
        const int mt = rest % m_tiles;
        const int sp = rest / m_tiles;
        const int m0 = mt * 256 + static_cast<int>(rank) * 128;       // this CTA's 128 rows of A / of the output
        const int n0 = nt * P_BN + static_cast<int>(rank) * 128;      // this CTA's half of the B tile
        const int kb0 = sp * k_per_split;
        const int kb1 = min(kb0 + k_per_split, k_blocks_total);
        if (p.m_active != nullptr && p.m_active[mt] == 0) { tile = tile_next; continue; }   // padding-only row tile
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!kb_active(p.k_active, kb)) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);   // the leader's barrier
          uint8_t* sa = smem + stage * P_STAGE_BYTES;
          uint8_t* sb = sa + P_A_BYTES;
          if (!A_MN) {
            tma_load_2d_pair(sa, &tmA, fb, kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_pair(sa + c * 8192, &tmA, fb, m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, &tmB, fb, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_pair(sb + c * 8192, &tmB, fb, n0 + c * 64, kb * BK);
          }
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
        if (XK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem + stage * P_STAGE_BYTES;
          uint8_t* sb = sa + P_A_BYTES;
          const int xb = (mt * 256) / p.xk_rows;                       // the sequence this 256-row tile lies in
          tma_load_2d_pair(sa, &tmA2, fb, 0, m0);
#pragma unroll
          for (int c = 0; c < 2; ++c) tma_load_2d_pair(sb + c * 8192, &tmB2, fb, n0 + c * 64, xb * 64);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
        tile = tile_next;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    // The whole warp stays converged; descriptors are warp-uniform values (base + stage / k offsets in
    // the 16-byte-granular address field) and only the tcgen05 instructions themselves are predicated on
    // the elected lane, which keeps the per-k-block issue sequence short.
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, P_BN, A_MN, B_MN);
      const bool elected = elect_one();
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t adesc0 = A_MN ? umma_smem_desc(smem_base, 8192, 1024) : umma_smem_desc(smem_base, 16, 1024);
      const uint64_t bdesc0 = B_MN ? umma_smem_desc(smem_base + P_A_BYTES, 8192, 1024)
                                   : umma_smem_desc(smem_base + P_A_BYTES, 16, 1024);
      constexpr uint32_t A_KSTEP = (A_MN ? 2048u : 32u) >> 4, B_KSTEP = (B_MN ? 2048u : 32u) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (;;) {
        const int tile = read_tile();
        __syncwarp();
        if (lane == 0 && tile >= 0) release_tile();
        advance_ring();
        if (tile >= total_tiles) break;
        const int sp = (tile / n_tiles) / m_tiles;
        const int kb0 = sp * k_per_split;
        const int kb1 = min(kb0 + k_per_split, k_blocks_total);
        if (p.m_active != nullptr && p.m_active[(tile / n_tiles) % m_tiles] == 0) continue;
        const int nkb = count_active_kb(p.k_active, kb0, kb1) + (XK ? 1 : 0);
        if (nkb == 0) continue;                       // wgrad slice made of padding only: nothing to add
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * P_BN;
        for (int it = 0; it < nkb; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t soff = static_cast<uint32_t>(stage) * (P_STAGE_BYTES >> 4);
          const uint64_t ad = adesc0 + soff, bd = bdesc0 + soff;
          if (elected) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(d_tmem, ad + k * A_KSTEP, bd + k * B_KSTEP, idesc, (it > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (it == nkb - 1) umma_commit_pair(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps (both CTAs, each on its own 128 accumulator rows) =================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;
    constexpr int CH = P_BN / 2 / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (;;) {
      const int tile = read_tile();
      __syncwarp();
      if (lane == 0 && tile >= 0) release_tile();
      advance_ring();
      if (tile >= total_tiles) break;
      const int nt = tile % n_tiles;
      const int mt = (tile / n_tiles) % m_tiles;
      if (p.m_active != nullptr && p.m_active[mt] == 0) continue;
      if (p.k_active != nullptr) {
        const int sp = (tile / n_tiles) / m_tiles;
        const int kb0 = sp * k_per_split;
        if (count_active_kb(p.k_active, kb0, min(kb0 + k_per_split, k_blocks_total)) == 0) continue;
      }
      const int m0 = mt * 256 + static_cast<int>(rank) * 128, n0 = nt * P_BN;
      const int row_base = m0 + quad * 32;
      uint8_t* my_stage = s_stage + (warp - 2) * 4096;
      float* sb = s_bias + acc * P_BN;
      if (etid < P_BN) sb[etid] = (p.bias != nullptr && n0 + etid < p.N) ? __ldg(p.bias + n0 + etid) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * P_BN + half * (P_BN / 2);
      uint32_t rbuf[2][32];
      tmem_ld32(tbase, rbuf[0]);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        tmem_ld_wait();
        if (c + 1 < CH) {
          tmem_ld32(tbase + (c + 1) * 32, rbuf[(c + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // leader's barrier
        }
        const int lcol = half * (P_BN / 2) + c * 32;
        const int col0 = n0 + lcol;
        if (col0 >= p.N || p.debug_skip_epilogue) continue;
        epilogue_chunk<EPI, OUT_F32, RES, DROP, SPLITK>(rbuf[c & 1], p, sb, my_stage, lane, lcol, row_base, col0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();     // neither CTA may exit (or free TMEM) while its peer still uses its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// One of the 256 scheduler counters of the current device, round-robin: launches that could be in flight at the
// same time (different streams) get different words; each launch leaves its word at zero.
static int* next_tile_counter() {
  static int* base[64] = {nullptr};
  static std::atomic<unsigned> seq{0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (base[dev] == nullptr) {
    void* ptr = nullptr;
    if (cudaGetSymbolAddress(&ptr, g_gemm_tile_counters) != cudaSuccess) {
      set_error(RF_ERR_CUDA, "rf_gemm_bf16: cudaGetSymbolAddress(g_gemm_tile_counters) failed");
      return nullptr;
    }
    base[dev] = reinterpret_cast<int*>(ptr);
  }
  return base[dev] + (seq.fetch_add(1, std::memory_order_relaxed) & 255u);
}

template <bool A_MN, bool B_MN, int EPI, bool OUT_F32, int RES, bool DROP, bool SPLITK, bool XK = false>
static int launch_gemm_pair(const rf_gemm_args* a, cudaStream_t stream) {
  auto kern = gemm_pair_kernel<A_MN, B_MN, EPI, OUT_F32, RES, DROP, SPLITK, XK>;
  static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
  if (first_use_on_device(&attr_seen)) {
    RF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
  }
  const CUtensorMap* tmA = A_MN ? get_tmap_2d(a->A, a->K, a->M, a->lda, 64) : get_tmap_2d(a->A, a->M, a->K, a->lda, 128);
  if (!tmA) return RF_ERR_CUDA;
  const CUtensorMap* tmB = B_MN ? get_tmap_2d(a->B, a->K, a->N, a->ldb, 64) : get_tmap_2d(a->B, a->N, a->K, a->ldb, 128);
  if (!tmB) return RF_ERR_CUDA;
  const CUtensorMap *tmA2 = tmA, *tmB2 = tmB;
  if (XK) {
    tmA2 = get_tmap_2d(a->A2, a->M, 64, 64, 128);
    if (!tmA2) return RF_ERR_CUDA;
    tmB2 = get_tmap_2d(a->B2, (a->M / a->xk_rows) * 64, a->N, a->N, 64);
    if (!tmB2) return RF_ERR_CUDA;
  }
  GemmParams p;
  fill_params(a, p);
  p.xk_rows = XK ? a->xk_rows : 0;
  const int m_tiles = (a->M + 255) / 256, n_tiles = (a->N + P_BN - 1) / P_BN;
  const int total = m_tiles * n_tiles * p.split_k;
  const int pairs = total < sm_count() / 2 ? total : sm_count() / 2;
  {
    const RowActivity& ra = row_activity();
    const bool wgrad_layout = A_MN && B_MN;
    p.m_active = (ra.flags != nullptr && !wgrad_layout && ra.rows == a->M) ? ra.flags : nullptr;
    p.k_active = (ra.flags != nullptr && wgrad_layout && ra.rows == a->K) ? ra.flags : nullptr;
  }
  static const bool static_sched = getenv("RF_GEMM_STATIC_SCHEDULE") != nullptr;
  p.tile_counter = static_sched ? nullptr : next_tile_counter();
  if (!static_sched && !p.tile_counter) return RF_ERR_CUDA;
  kern<<<2 * pairs, GEMM_THREADS, P_SMEM, stream>>>(*tmA, *tmB, *tmA2, *tmB2, p);
  return check_launch("rf_gemm_bf16(pair)");
}

template <int BN, bool A_MN, bool B_MN, int EPI, bool OUT_F32, int RES, bool DROP, bool SPLITK>
static int launch_gemm(const rf_gemm_args* a, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI, OUT_F32, RES, DROP, SPLITK>;
  static std::atomic<unsigned long long> attr_seen{0};   // one bit per device
  if (first_use_on_device(&attr_seen)) {
    RF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  }
  const CUtensorMap* tmA = A_MN ? get_tmap_2d(a->A, a->K, a->M, a->lda, 64) : get_tmap_2d(a->A, a->M, a->K, a->lda, BM);
  if (!tmA) return RF_ERR_CUDA;
  const CUtensorMap* tmB = B_MN ? get_tmap_2d(a->B, a->K, a->N, a->ldb, 64) : get_tmap_2d(a->B, a->N, a->K, a->ldb, BN);
  if (!tmB) return RF_ERR_CUDA;
  GemmParams p;
  fill_params(a, p);
  p.m_active = p.k_active = nullptr;       // the single-CTA kernel (small problems) takes no row activity
  p.tile_counter = nullptr;
  const int m_tiles = (a->M + BM - 1) / BM, n_tiles = (a->N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * p.split_k;
  const int grid = total < sm_count() ? total : sm_count();
  kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(*tmA, *tmB, p);
  return check_launch("rf_gemm_bf16");
}

}  // namespace rf

extern "C" int rf_gemm_bf16(const rf_gemm_args* a, rf_stream_t stream_) {
  using namespace rf;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(a != nullptr, "rf_gemm_bf16: null args");
  RF_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "rf_gemm_bf16: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  RF_REQUIRE(a->N % 32 == 0, "rf_gemm_bf16: N=%d must be a multiple of 32", a->N);
  RF_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "rf_gemm_bf16: lda/ldb must be multiples of 8 (TMA 16B strides)");
  RF_REQUIRE(a->ldc % 8 == 0, "rf_gemm_bf16: ldc must be a multiple of 8");
  RF_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->C) & 15) == 0,
             "rf_gemm_bf16: A/B/C must be 16-byte aligned");
  RF_REQUIRE(a->scale_ncols % 32 == 0, "rf_gemm_bf16: scale_ncols must be a multiple of 32");
  RF_REQUIRE(a->epi == RF_EPI_NONE || a->scale_ncols == 0, "rf_gemm_bf16: the GELU / dGELU epilogues take no column scale");
  RF_REQUIRE(a->split_k <= 1 || a->out_f32, "rf_gemm_bf16: split_k needs fp32 output");
  RF_REQUIRE(a->epi != RF_EPI_GELU || (a->C2 != nullptr && !a->out_f32), "rf_gemm_bf16: GELU epilogue needs C2, bf16");
  RF_REQUIRE(a->epi != RF_EPI_DGELU || a->aux != nullptr, "rf_gemm_bf16: DGELU epilogue needs aux");
  RF_REQUIRE(a->A2 == nullptr || (!a->a_mn_major && a->b_mn_major),
             "rf_gemm_bf16: the extra k-block (A2/B2) is instantiated for the dgrad layout only");
  const int layout = (a->a_mn_major ? 2 : 0) | (a->b_mn_major ? 1 : 0);
  // CTA-pair kernel (256 x 256 tiles) whenever there is more than one 128-row slab of output;
  // the single-CTA kernel covers small problems (e.g. one short sequence).
  const bool pair = a->M > 128 && a->N >= 128;
  // what the (branch-free) epilogue has to do, resolved to template flags
  const bool drop = a->drop_p > 0.f;
  int split = 1;
  {
    const int kb = (a->K + BK - 1) / BK;
    int sk = a->split_k > 1 ? a->split_k : 1;
    if (sk > kb) sk = kb;
    const int per = (kb + sk - 1) / sk;
    split = (kb + per - 1) / per;
  }
  const bool splitk = split > 1;
  const bool acc_c = a->accumulate && a->out_f32 && a->residual == nullptr;
  RF_REQUIRE(!(a->accumulate && a->residual != nullptr), "rf_gemm_bf16: accumulate and residual are exclusive");
  const int res = (a->residual != nullptr) ? (a->residual_f32 ? 2 : 1) : ((acc_c && !splitk) ? 2 : 0);
#define RF_GEMM(BNS, AMN, BMN, EPIK, F32, RESK, DROPK, SPLK)                                          \
  (pair ? launch_gemm_pair<AMN, BMN, EPIK, F32, RESK, DROPK, SPLK>(a, stream)                          \
        : launch_gemm<BNS, AMN, BMN, EPIK, F32, RESK, DROPK, SPLK>(a, stream))
  if (layout == 0) {
    if (a->epi == RF_EPI_GELU) {
      RF_REQUIRE(res == 0 && !drop && !splitk, "rf_gemm_bf16: the GELU epilogue takes neither residual nor dropout");
      return RF_GEMM(256, false, false, RF_EPI_GELU, false, 0, false, false);
    }
    RF_REQUIRE(a->epi == RF_EPI_NONE, "rf_gemm_bf16: epilogue %d unsupported for K-major x K-major", a->epi);
    RF_REQUIRE(!splitk, "rf_gemm_bf16: split_k is only instantiated for the wgrad layout");
    if (a->out_f32) {
      RF_REQUIRE(res != 1, "rf_gemm_bf16: fp32 output takes an fp32 residual");
      if (res == 2) return drop ? RF_GEMM(128, false, false, RF_EPI_NONE, true, 2, true, false)
                                : RF_GEMM(128, false, false, RF_EPI_NONE, true, 2, false, false);
      return drop ? RF_GEMM(128, false, false, RF_EPI_NONE, true, 0, true, false)
                  : RF_GEMM(128, false, false, RF_EPI_NONE, true, 0, false, false);
    }
    RF_REQUIRE(res != 2 && !drop, "rf_gemm_bf16: bf16 output takes a bf16 residual and no dropout in the forward layout");
    return res == 1 ? RF_GEMM(256, false, false, RF_EPI_NONE, false, 1, false, false)
                    : RF_GEMM(256, false, false, RF_EPI_NONE, false, 0, false, false);
  }
  if (layout == 1) {  // dgrad: dY[M,K] (K-major) x W stored [K,N]
    RF_REQUIRE(!a->out_f32 && !drop && !splitk, "rf_gemm_bf16: the dgrad layout writes bf16 without dropout / split-K");
    if (a->epi == RF_EPI_DGELU) {
      RF_REQUIRE(res == 0, "rf_gemm_bf16: the dGELU epilogue takes no residual");
      return RF_GEMM(256, false, true, RF_EPI_DGELU, false, 0, false, false);
    }
    RF_REQUIRE(a->epi == RF_EPI_NONE, "rf_gemm_bf16: epilogue %d unsupported for the dgrad layout", a->epi);
    RF_REQUIRE(res != 2, "rf_gemm_bf16: the dgrad layout takes a bf16 residual");
    if (a->A2 != nullptr) {
      RF_REQUIRE(a->B2 != nullptr && pair && res == 1 && a->xk_rows > 0 && a->xk_rows % 256 == 0 &&
                     a->M % a->xk_rows == 0 && a->N % 8 == 0 &&
                     ((reinterpret_cast<uintptr_t>(a->A2) | reinterpret_cast<uintptr_t>(a->B2)) & 15) == 0,
                 "rf_gemm_bf16: the extra k-block needs a bf16 residual, 16-byte aligned A2/B2 and xk_rows %% 256 == 0 "
                 "dividing M (got M=%d xk_rows=%d)", a->M, a->xk_rows);
      return launch_gemm_pair<false, true, RF_EPI_NONE, false, 1, false, false, true>(a, stream);
    }
    return res == 1 ? RF_GEMM(256, false, true, RF_EPI_NONE, false, 1, false, false)
                    : RF_GEMM(256, false, true, RF_EPI_NONE, false, 0, false, false);
  }
  if (layout == 3) {  // wgrad: dY^T x X, both stored [K, *]
    RF_REQUIRE(a->out_f32 && a->epi == RF_EPI_NONE && !drop && a->residual == nullptr,
               "rf_gemm_bf16: the wgrad layout writes fp32 without epilogue");
    if (splitk) return RF_GEMM(128, true, true, RF_EPI_NONE, true, 0, false, true);
    return res == 2 ? RF_GEMM(128, true, true, RF_EPI_NONE, true, 2, false, false)
                    : RF_GEMM(128, true, true, RF_EPI_NONE, true, 0, false, false);
  }
#undef RF_GEMM
  return set_error(RF_ERR_INVALID, "rf_gemm_bf16: layout a_mn_major=1,b_mn_major=0 is not instantiated");
}
