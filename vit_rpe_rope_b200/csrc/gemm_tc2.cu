// bf16 GEMMs of the ViT hot path on tcgen05 tensor cores with CTA PAIRS (cta_group::2), all operand
// layouts, fused epilogues.  One kernel template serves
//   * the QKV projection with the RoPE rotate-half epilogue         (models/vit.py:47-68)
//   * the patch-embedding GEMM                                      (models/vit.py:164,248-254)
//   * out-proj / fc1(+bias+GELU) / fc2 / head forward               (models/vit.py:91,118,285)
//   * every projection backward: dX = dY . W  and  dW = dY^T . X    (autograd of the above)
//
// C[M][N] = op(A)[M][K] . op(B)[K][N], fp32 accumulation in TMEM.
//   A_MN = 0: A stored [M][K] (K-major);  A_MN = 1: A stored [K][M] (MN-major, i.e. A^T is given)
//   B_MN = 0: B stored [N][K] (K-major, a Linear weight); B_MN = 1: B stored [K][N]
//
// A cluster of two CTAs (one per SM of a TPC) computes a 256 x BN tile: each CTA stages ITS 128 rows of
// op(A) and ITS half (BN/2 columns) of op(B) with TMA (128-byte swizzle), the leader CTA's MMA thread
// issues tcgen05.mma.cta_group::2 (M 256, N BN, K 16) which reads both CTAs' shared memory, and each CTA's
// TMEM receives its 128 rows x BN columns.  Per SM and per MMA this halves the shared-memory operand
// traffic (8 KB read per 128 tensor-clocks instead of 12 KB) - the 1-CTA kernel of round 1 was bound by
// shared-memory bandwidth (TMA fill + operand reads) at ~45 % of the tensor peak.
//
// Persistent: gridDim.x / 2 clusters loop over work units (m-tile, n-tile, k-split); warp roles
//   warp 0        TMA producer (one lane), kStages-deep ring, completes on the LEADER's full barrier
//   warp 1        TMEM allocator (both CTAs) + MMA issuer (leader only); accumulators double-buffered
//   warps 2..9    epilogue: tcgen05.ld -> registers -> (bias / GELU / RoPE) -> swizzled shared staging
//                 -> TMA store / TMA reduce-add (fp32 split-K), or a direct mapped store (planes, tokens)
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrr {

using namespace tc;

namespace {

constexpr int kPM = 128;                              // rows of op(A) per CTA
constexpr int kPK = 64;                               // K elements per pipeline stage
constexpr int kEpiWarps = 8;
constexpr int kPairThreads = 64 + 32 * kEpiWarps;     // 320
constexpr int kWarpStaging = 8192;                    // two [32 rows][128 B] swizzled buffers per epilogue warp
constexpr int kSmemBudget = 232448;                   // 227 KB

struct PairShape {
  int M, N, K;
  int tiles_m, tiles_n, splits, num_k;
};

// ---------------------------------------------------------------------------------------------------
// Epilogues.  kTma = false: functor called per (row, 64-column chunk) with the fp32 accumulators.
// ---------------------------------------------------------------------------------------------------
// kPlaneStore: the functor only transforms the 64 accumulators of (row m, head chunk n0) in registers; the kernel
// rounds them to bf16, stages 32 rows x 128 B per warp in swizzled shared memory and writes them with TMA through a
// 3-D map of planes[3 B H][N][64]; the one warp tile in N / 32 that runs into the next image falls back to per-thread
// row stores.  A thread writing its own 128-byte row touched 32 lines per store instruction and made
// this GEMM 1.7x slower than the plain one.
struct QkvRopeEpi2 {
  static constexpr bool kTma = false;
  static constexpr bool kPlaneStore = true;
  __nv_bfloat16* planes;
  const float *cos_tab, *sin_tab;
  const float4* packed;  // rope_pack_tables layout [heads][16][N-1], or null
  int B, N, E, H, rope_mode;
  __device__ __forceinline__ void rotate(int m, int n0, float (&f)[64]) const {
    const int hd = 32;
    const int b = m / N, t = m - b * N;
    const int which = n0 / E, h = (n0 - which * E) >> 6;
    if (rope_mode != VRR_ROPE_NONE && which < 2 && t >= 1 && b < B && packed != nullptr) {
      // lanes = consecutive token rows = consecutive float4 of the packed table: coalesced 512-byte requests
      const float4* pk = packed + (size_t)(rope_mode == VRR_ROPE_MIXED ? h * 16 : 0) * (N - 1) + (t - 1);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 v = __ldg(pk + (size_t)q * (N - 1));  // {cos[2q], sin[2q], cos[2q+1], sin[2q+1]}
        const float x1 = f[2 * q], x2 = f[2 * q + hd], y1 = f[2 * q + 1], y2 = f[2 * q + 1 + hd];
        f[2 * q] = x1 * v.x - x2 * v.y;
        f[2 * q + hd] = x1 * v.y + x2 * v.x;
        f[2 * q + 1] = y1 * v.z - y2 * v.w;
        f[2 * q + 1 + hd] = y1 * v.w + y2 * v.z;
      }
    } else if (rope_mode != VRR_ROPE_NONE && which < 2 && t >= 1 && b < B) {
      const size_t base = ((size_t)(rope_mode == VRR_ROPE_MIXED ? h * (N - 1) : 0) + (t - 1)) * hd;
      const float4* c4 = reinterpret_cast<const float4*>(cos_tab + base);
      const float4* s4 = reinterpret_cast<const float4*>(sin_tab + base);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const float4 c = __ldg(c4 + q4), s = __ldg(s4 + q4);
        const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = q4 * 4 + e;
          const float x1 = f[d], x2 = f[d + hd];
          f[d] = x1 * cc[e] - x2 * ss[e];
          f[d + hd] = x1 * ss[e] + x2 * cc[e];
        }
      }
    }
  }
  // plane index (third map coordinate) of the head chunk n0 for image b
  __device__ __forceinline__ int plane_of(int b, int n0) const {
    const int which = n0 / E, h = (n0 - which * E) >> 6;
    return (which * B + b) * H + h;
  }
};

// tokens[b][1+p][n] = round_bf16(acc + bias[n]) (+ pos[p][n]); TT = token-stream type
template <typename TT>
struct PatchEmbedEpi2 {
  static constexpr bool kTma = false;
  static constexpr bool kPlaneStore = false;
  TT* tokens;
  const __nv_bfloat16* bias;
  const TT* pos;  // may be null
  int Np, E;
  __device__ __forceinline__ void operator()(int m, int n0, float (&f)[64]) const {
    const int b = m / Np, p = m - b * Np;
    TT* dst = tokens + ((size_t)b * (Np + 1) + 1 + p) * E + n0;
    const TT* prow = pos ? pos + (size_t)p * E + n0 : nullptr;
#pragma unroll
    for (int v4 = 0; v4 < 16; ++v4) {
      const float4 bv = ld4(bias + n0 + v4 * 4);
      float4 o;
      o.x = __bfloat162float(__float2bfloat16_rn(f[v4 * 4 + 0] + bv.x));
      o.y = __bfloat162float(__float2bfloat16_rn(f[v4 * 4 + 1] + bv.y));
      o.z = __bfloat162float(__float2bfloat16_rn(f[v4 * 4 + 2] + bv.z));
      o.w = __bfloat162float(__float2bfloat16_rn(f[v4 * 4 + 3] + bv.w));
      if (prow) {
        const float4 pv = ld4(prow + v4 * 4);
        o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
      }
      st4(dst + v4 * 4, o);
    }
  }
};

// Row-major C through shared staging + TMA store.  Runtime (warp-uniform) switches keep the number of
// template instances small.
struct StoreEpi {
  static constexpr bool kTma = true;
  static constexpr bool kPlaneStore = false;
  const float* bias;  // fp32 [N] or null; rounded to bf16 first when C is bf16 (autocast casts the bias)
  int out_f32;        // C element type: 0 bf16, 1 fp32
  int reduce;         // fp32 only: C += tile via TMA reduce-add (split-K partial sums / gradient accumulation)
  int gelu;           // bf16 only.  1: C = h = bf16(acc + bias), C2 = bf16(gelu(h)) (exact erf GELU);
                      //            2: C = bf16(gelu(h)), C2 = bf16(gelu'(h)) - what an Mlp backward needs instead of h
                      //            3: C = bf16(gelu(h)) only (inference): its own code path, one TMA store per chunk
  const __nv_bfloat16* mul;  // bf16 only, or null: C = bf16(acc * mul[m][n]) (mul is [M][N] like C): d_act * gelu'(h)
  float* colsum;      // bf16 plain / bias / mul paths, or null: colsum[n] += sum_m C[m][n] (values as stored; atomics)
};

// gelu(h) and d gelu / dh for TWO values at a time on the packed fp32x2 pipe (fma.rn.f32x2, sm_100): the epilogue of
// fc1 is bound by its instruction count (128 x 256 outputs per CTA against ~6 100 tensor-clocks of MMA per tile), and
// the scalar version (~20 FMA-pipe ops per element) made the GEMM run at half the speed of the plain one.
//   Phi(h) = 1 - erfc(|h| / sqrt 2) / 2 (h >= 0), erfc / 2 otherwise; erfc(z) = (a1 t + .. + a5 t^5) exp(-z^2),
//   t = 1 / (1 + p z) (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7); gelu = h Phi, gelu' = Phi + h phi(h) with
//   phi(h) = exp(-h^2 / 2) / sqrt(2 pi) - the same exponential.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_and_grad2(float2 h, float2& g, float2& dg) {
  const float2 hh = __fmul2_rn(h, h);
  const float2 ea = __fmul2_rn(hh, make_float2(-0.72134752044448170f, -0.72134752044448170f));
  const float2 e = make_float2(ex2(ea.x), ex2(ea.y));  // exp(-h^2 / 2)
  const float2 za = make_float2(fabsf(h.x), fabsf(h.y));
  const float2 den = __ffma2_rn(za, make_float2(0.23164190f, 0.23164190f), make_float2(1.f, 1.f));  // 1 + p |h| / sqrt 2
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  // coefficients pre-multiplied by 1/2
  float2 poly = __ffma2_rn(make_float2(0.5307027145f, 0.5307027145f), t, make_float2(-0.7265760135f, -0.7265760135f));
  poly = __ffma2_rn(poly, t, make_float2(0.7107068705f, 0.7107068705f));
  poly = __ffma2_rn(poly, t, make_float2(-0.142248368f, -0.142248368f));
  poly = __ffma2_rn(poly, t, make_float2(0.127414796f, 0.127414796f));
  const float2 q = __fmul2_rn(__fmul2_rn(poly, t), e);  // erfc(|h| / sqrt 2) / 2
  const float2 om = __ffma2_rn(q, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
  const float2 cdf = make_float2(h.x >= 0.f ? om.x : q.x, h.y >= 0.f ? om.y : q.y);  // Phi(h)
  g = __fmul2_rn(h, cdf);
  dg = __ffma2_rn(__fmul2_rn(h, e), make_float2(0.3989422804014327f, 0.3989422804014327f), cdf);
}

// gelu(h) alone, packed to bf16x2 (same Phi(h) as above)
__device__ __forceinline__ uint32_t gelu_only2(float2 h) {
  const float2 hh = __fmul2_rn(h, h);
  const float2 ea = __fmul2_rn(hh, make_float2(-0.72134752044448170f, -0.72134752044448170f));
  const float2 e = make_float2(ex2(ea.x), ex2(ea.y));
  const float2 za = make_float2(fabsf(h.x), fabsf(h.y));
  const float2 den = __ffma2_rn(za, make_float2(0.23164190f, 0.23164190f), make_float2(1.f, 1.f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 poly = __ffma2_rn(make_float2(0.5307027145f, 0.5307027145f), t, make_float2(-0.7265760135f, -0.7265760135f));
  poly = __ffma2_rn(poly, t, make_float2(0.7107068705f, 0.7107068705f));
  poly = __ffma2_rn(poly, t, make_float2(-0.142248368f, -0.142248368f));
  poly = __ffma2_rn(poly, t, make_float2(0.127414796f, 0.127414796f));
  const float2 q = __fmul2_rn(__fmul2_rn(poly, t), e);
  const float2 om = __ffma2_rn(q, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
  const float2 g = __fmul2_rn(h, make_float2(h.x >= 0.f ? om.x : q.x, h.y >= 0.f ? om.y : q.y));
  return pack_bf16(g.x, g.y);
}

// ---------------------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2,
                 const PairShape g, const Epi epi) {
  constexpr int kBHalf = BN / 2;
  constexpr int kABytes = kPM * kPK * 2;
  constexpr int kBBytes = kBHalf * kPK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kStages = (kSmemBudget - 1024 - kEpiWarps * kWarpStaging - 512) / kStageBytes;
  constexpr uint32_t kTmemCols = 2 * BN;  // two accumulator stages (256 or 512: powers of two)
  static_assert(kStages >= 3, "pipeline too shallow");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kEpiWarps * kWarpStaging);
  uint64_t* bar_full = bars;                              // [kStages] leader's: both CTAs' TMA -> MMA
  uint64_t* bar_empty = bars + kStages;                   // [kStages] per CTA: MMA retired -> own TMA producer
  uint64_t* bar_acc_full = bars + 2 * kStages;            // [2] per CTA: accumulator ready -> own epilogue
  uint64_t* bar_acc_empty = bars + 2 * kStages + 2;       // [2] leader's: both CTAs' epilogues -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int total_units = g.tiles_m * g.tiles_n * g.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], 2);   // one arrive.expect_tx per CTA of the pair
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_acc_full[a], 1);
      mbar_init(&bar_acc_empty[a], 2 * kEpiWarps);  // one elected lane per epilogue warp, both CTAs
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync();  // both CTAs: barriers initialised, TMEM allocated, before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ====================================== TMA producer (both CTAs) ======================================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_a);
      tma_prefetch_desc(&tmap_b);
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
        const int split = unit % g.splits, t = unit / g.splits;
        const int n_blk = t % g.tiles_n, m_blk = t / g.tiles_n;
        const int kb0 = (int)((long long)split * g.num_k / g.splits), kb1 = (int)((long long)(split + 1) * g.num_k / g.splits);
        const int m_row0 = m_blk * (2 * kPM) + (int)rank * kPM;
        const int n_row0 = n_blk * BN + (int)rank * kBHalf;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_cluster(&bar_empty[stage], phase ^ 1);
          const uint32_t full_leader = mapa(smem_u32(&bar_full[stage]), 0);
          mbar_expect_tx_cluster(full_leader, kStageBytes);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (!A_MN) {
            tma_load_2d_pair(sa, &tmap_a, full_leader, kb * kPK, m_row0);
          } else {
#pragma unroll
            for (int j = 0; j < kPM / 64; ++j) tma_load_2d_pair(sa + j * 8192, &tmap_a, full_leader, m_row0 + 64 * j, kb * kPK);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, &tmap_b, full_leader, kb * kPK, n_row0);
          } else {
#pragma unroll
            for (int j = 0; j < kBHalf / 64; ++j) tma_load_2d_pair(sb + j * 8192, &tmap_b, full_leader, n_row0 + 64 * j, kb * kPK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer (leader CTA only) ==================================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(2 * kPM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t adv_a = A_MN ? 128 : 2, adv_b = B_MN ? 128 : 2;  // descriptor units (16 B) per K = 16
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
        const int split = unit % g.splits;
        const int kb0 = (int)((long long)split * g.num_k / g.splits), kb1 = (int)((long long)(split + 1) * g.num_k / g.splits);
        mbar_wait_cluster(&bar_acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_cluster(&bar_full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = A_MN ? smem_desc_sw128_ex(sa, 8192, 1024) : smem_desc_sw128(sa);
          const uint64_t db = B_MN ? smem_desc_sw128_ex(sa + kABytes, 8192, 1024) : smem_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kPK / 16; ++k)
            mma_ss_pair(d_tmem, da + adv_a * k, db + adv_b * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          mma_commit_pair(&bar_empty[stage], 3);  // both CTAs' producers may refill the slot when these retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        mma_commit_pair(&bar_acc_full[acc], 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ====================================== epilogue (both CTAs) ==========================================
    const int ew = warp - 2;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int chalf = ew >> 2;          // which half of the tile's columns
    constexpr int kChunks = BN / 2 / 64;
    uint8_t* my_stage = staging + ew * kWarpStaging;
    const uint32_t stage_u32 = smem_u32(my_stage);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_counter = 0;
    for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
      const int t = unit / g.splits;
      const bool first_split = (unit - t * g.splits) == 0;  // the bias joins exactly one of the k-splits
      const int n_blk = t % g.tiles_n, m_blk = t / g.tiles_n;
      const int m_warp0 = m_blk * (2 * kPM) + (int)rank * kPM + quad * 32;
      const int m = m_warp0 + lane;
      mbar_wait_cluster(&bar_acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + chalf * (BN / 2);
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[64];
        // StoreEpi::mul: the warp's [32 rows][64] multiplier tile.  Loaded COALESCED (instruction i = rows 4i..4i+3,
        // eight lanes x 16 bytes per row; a lane reading its own row touched 32 lines per instruction) while the TMEM
        // load is in flight, then turned into "lane = row" through the swizzled staging buffer this chunk's output
        // will use (its previous TMA store has been read: bulk_wait_read<1>).
        [[maybe_unused]] uint4 mv[8];
        [[maybe_unused]] bool with_mul = false;
        if constexpr (Epi::kTma) {
          with_mul = epi.mul != nullptr;
          if (with_mul) {
            const int n0m = n_blk * BN + chalf * (BN / 2) + c * 64;
            const int nb = max(min(n0m + (lane & 7) * 8, g.N - 8), 0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mr = min(m_warp0 + i * 4 + (lane >> 3), g.M - 1);
              mv[i] = __ldg(reinterpret_cast<const uint4*>(epi.mul + (size_t)mr * g.N + nb));
            }
          }
        }
        __syncwarp();  // tcgen05.ld is .sync.aligned
        tmem_ld32(trow + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(trow + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        if constexpr (Epi::kTma) {
          if (with_mul) {
            const uint32_t tbuf = stage_u32 + (chunk_counter & 1u) * 4096u;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t r = (uint32_t)(i * 4 + (lane >> 3));
              st_shared_v4(tbuf + r * 128u + ((((uint32_t)lane & 7u) ^ (r & 7u)) << 4), mv[i].x, mv[i].y, mv[i].z, mv[i].w);
            }
            __syncwarp();
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8)
              mv[c8] = ld_shared_v4(tbuf + (uint32_t)lane * 128u + ((((uint32_t)c8) ^ ((uint32_t)lane & 7u)) << 4));
          }
        }
        tmem_wait_ld();
        if (c == kChunks - 1) {  // this warp has read everything it needs from the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa(smem_u32(&bar_acc_empty[acc]), 0));
        }
        const int n0 = n_blk * BN + chalf * (BN / 2) + c * 64;
        float(&f)[64] = *reinterpret_cast<float(*)[64]>(v);
        if constexpr (Epi::kPlaneStore) {
          epi.rotate(m, n0, f);
          const int b0 = m_warp0 / epi.N, t0 = m_warp0 - b0 * epi.N;
          const bool in_range = n0 < g.N && m_warp0 < g.M;
          if (t0 + 32 <= epi.N || !in_range) {
            // the warp's 32 rows lie in one image: swizzled staging + one TMA store
            const uint32_t buf = (chunk_counter & 1u) * 4096u;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            const uint32_t sw = (uint32_t)(lane & 7);
            const uint32_t row_base = stage_u32 + buf + (uint32_t)lane * 128u;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8)
              st_shared_v4(row_base + (((uint32_t)c8 ^ sw) << 4), pack_bf16(f[c8 * 8 + 0], f[c8 * 8 + 1]),
                           pack_bf16(f[c8 * 8 + 2], f[c8 * 8 + 3]), pack_bf16(f[c8 * 8 + 4], f[c8 * 8 + 5]),
                           pack_bf16(f[c8 * 8 + 6], f[c8 * 8 + 7]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (in_range) tma_store_3d(&tmap_c, my_stage + buf, 0, t0, epi.plane_of(b0, n0));
              bulk_commit();
            }
            ++chunk_counter;
          } else if (m < g.M) {
            // the rows run into the next image (one warp tile in N / 32; TMA stores take no negative coordinates -
            // "illegal instruction" - and a box cannot be shortened): each thread stores its own row
            const int b = m / epi.N, t = m - b * epi.N;
            __nv_bfloat16* dst = epi.planes + ((size_t)epi.plane_of(b, n0) * epi.N + t) * 64;
#pragma unroll
            for (int v8 = 0; v8 < 8; ++v8) {
              uint4 wv;
              wv.x = pack_bf16(f[v8 * 8 + 0], f[v8 * 8 + 1]);
              wv.y = pack_bf16(f[v8 * 8 + 2], f[v8 * 8 + 3]);
              wv.z = pack_bf16(f[v8 * 8 + 4], f[v8 * 8 + 5]);
              wv.w = pack_bf16(f[v8 * 8 + 6], f[v8 * 8 + 7]);
              *reinterpret_cast<uint4*>(dst + v8 * 8) = wv;
            }
          }
        } else if constexpr (!Epi::kTma) {
          if (m < g.M && n0 < g.N) epi(m, n0, f);
        } else {
          const bool store_ok = (n0 < g.N) && (m_warp0 < g.M);
          const uint32_t sw = (uint32_t)(lane & 7);
          const uint32_t row_base = stage_u32 + (uint32_t)lane * 128u;
          if (!epi.out_f32) {
            // ---- bf16 C ----
            if (epi.gelu == 3) {
              // inference: gelu(h) alone.  Its own branch - folding it into the two-output code below cost that path
              // 50 us per launch (register allocation of the o0 / o1 arrays).
              uint32_t o0[32];
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                const int nb = min(n0 + c8 * 8, g.N - 8);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + nb));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + nb + 4));
                const uint32_t br[4] = {pack_bf16(b0.x, b0.y), pack_bf16(b0.z, b0.w), pack_bf16(b1.x, b1.y), pack_bf16(b1.z, b1.w)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 bb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&br[j]));
                  const float2 hs = __fadd2_rn(make_float2(f[c8 * 8 + 2 * j], f[c8 * 8 + 2 * j + 1]), bb);
                  const uint32_t hw = pack_bf16(hs.x, hs.y);
                  o0[c8 * 4 + j] = gelu_only2(__bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw)));
                }
              }
              const uint32_t buf = (chunk_counter & 1u) * 4096u;
              if (lane == 0) bulk_wait_read<1>();
              __syncwarp();
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8)
                st_shared_v4(row_base + buf + (((uint32_t)c8 ^ sw) << 4), o0[c8 * 4], o0[c8 * 4 + 1], o0[c8 * 4 + 2],
                             o0[c8 * 4 + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (store_ok) tma_store_2d(&tmap_c, my_stage + buf, n0, m_warp0);
                bulk_commit();
              }
            } else if (epi.gelu != 0) {
              // two outputs: (h, gelu(h)) or (gelu(h), gelu'(h)).  Computed and flushed to the staging buffers in two
              // halves of 32 columns: the profile of the one-piece version was 58 % fixed-latency dependency stalls
              // (`stall_wait`) - with 64 packed outputs held next to the 64 accumulators ptxas had no registers left
              // to interleave the per-pair chains.  The first half is computed BEFORE the wait for the previous
              // chunk's TMA stores.
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                uint32_t o0[16], o1[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int c8 = hf * 4 + q;
                  const int nb = min(n0 + c8 * 8, g.N - 8);  // N % 8 == 0; clamp keeps the tail read in bounds
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + nb));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + nb + 4));
                  const uint32_t br[4] = {pack_bf16(b0.x, b0.y), pack_bf16(b0.z, b0.w), pack_bf16(b1.x, b1.y), pack_bf16(b1.z, b1.w)};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 bb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&br[j]));  // bias as bf16
                    const float2 hs = __fadd2_rn(make_float2(f[c8 * 8 + 2 * j], f[c8 * 8 + 2 * j + 1]), bb);
                    const uint32_t hw = pack_bf16(hs.x, hs.y);
                    const float2 hr = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw));  // h as stored
                    float2 gl, dg;
                    gelu_and_grad2(hr, gl, dg);
                    if (epi.gelu == 1) {
                      o0[q * 4 + j] = hw;
                      o1[q * 4 + j] = pack_bf16(gl.x, gl.y);
                    } else {
                      o0[q * 4 + j] = pack_bf16(gl.x, gl.y);
                      o1[q * 4 + j] = pack_bf16(dg.x, dg.y);
                    }
                  }
                }
                if (hf == 0) {
                  if (lane == 0) bulk_wait_read<0>();
                  __syncwarp();
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t off = (((uint32_t)(hf * 4 + q) ^ sw) << 4);
                  st_shared_v4(row_base + off, o0[q * 4], o0[q * 4 + 1], o0[q * 4 + 2], o0[q * 4 + 3]);
                  st_shared_v4(row_base + 4096u + off, o1[q * 4], o1[q * 4 + 1], o1[q * 4 + 2], o1[q * 4 + 3]);
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (store_ok) {
                  tma_store_2d(&tmap_c, my_stage, n0, m_warp0);
                  tma_store_2d(&tmap_c2, my_stage + 4096, n0, m_warp0);
                }
                bulk_commit();
              }
            } else {
              const uint32_t buf = (chunk_counter & 1u) * 4096u;
              if (lane == 0) bulk_wait_read<1>();
              __syncwarp();
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = f[c8 * 8 + j];
                if (epi.bias != nullptr) {
                  const int nb = min(n0 + c8 * 8, g.N - 8);
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + nb));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + nb + 4));
                  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) x[j] += __bfloat162float(__float2bfloat16_rn(bb[j]));
                }
                if (epi.mul != nullptr) {
                  const uint32_t mw[4] = {mv[c8].x, mv[c8].y, mv[c8].z, mv[c8].w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 m2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&mw[j]));
                    x[2 * j] *= m2.x;
                    x[2 * j + 1] *= m2.y;
                  }
                }
                st_shared_v4(row_base + buf + (((uint32_t)c8 ^ sw) << 4), pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                             pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (store_ok) tma_store_2d(&tmap_c, my_stage + buf, n0, m_warp0);
                bulk_commit();
              }
              if (epi.colsum != nullptr && store_ok) {
                // bias gradient of the layer that produced A: column sums of the tile AS STORED (bf16), read back from
                // the staging buffer while the TMA store drains it; lane = column pair, conflict-free under the swizzle
                // (rows past M hold zeros: their A rows were zero-filled)
                float s0 = 0.f, s1 = 0.f;
                const uint32_t cbase = stage_u32 + buf + ((uint32_t)lane & 3u) * 4u;
#pragma unroll 8
                for (uint32_t r = 0; r < 32; ++r) {
                  const uint32_t wv = ld_shared_u32(cbase + r * 128u + ((((uint32_t)lane >> 2) ^ (r & 7u)) << 4));
                  const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv));
                  s0 += v2.x;
                  s1 += v2.y;
                }
                const int col = n0 + 2 * lane;
                if (col < g.N) {
                  atomicAdd(epi.colsum + col, s0);
                  atomicAdd(epi.colsum + col + 1, s1);
                }
              }
            }
          } else {
            // ---- fp32 C: two [32 rows][32 floats] boxes per 64-column chunk ----
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                float x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = f[hb * 32 + c4 * 4 + j];
                if (epi.bias != nullptr && first_split) {
                  const int nb = min(n0 + hb * 32 + c4 * 4, g.N - 4);
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + nb));
                  x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                }
                st_shared_v4(row_base + hb * 4096u + (((uint32_t)c4 ^ sw) << 4), __float_as_uint(x[0]), __float_as_uint(x[1]),
                             __float_as_uint(x[2]), __float_as_uint(x[3]));
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (store_ok) {
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                  if (n0 + hb * 32 < g.N) {
                    if (epi.reduce) tma_reduce_add_2d(&tmap_c, my_stage + hb * 4096, n0 + hb * 32, m_warp0);
                    else tma_store_2d(&tmap_c, my_stage + hb * 4096, n0 + hb * 32, m_warp0);
                  }
                }
              }
              bulk_commit();
            }
          }
          ++chunk_counter;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (Epi::kTma || Epi::kPlaneStore) {
      if (lane == 0) bulk_wait<0>();  // all stores of this warp have been written
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync();  // nobody signals the peer's barriers / reads its shared memory after this point
  if (warp == 1) tmem_dealloc_pair(tmem_base, kTmemCols);
}

template <int BN>
constexpr size_t pair_smem_bytes() {
  constexpr int stage = kPM * kPK * 2 + (BN / 2) * kPK * 2;
  constexpr int stages = (kSmemBudget - 1024 - kEpiWarps * kWarpStaging - 512) / stage;
  return 1024 + (size_t)stages * stage + kEpiWarps * kWarpStaging + 512;
}

// Pick the k-split count for accumulate/reduce GEMMs so that the work units fill whole waves of clusters.
int pick_splits(int tiles, int num_k, int clusters) {
  if (tiles >= 2 * clusters) return 1;
  int best = 1;
  double best_cost = 1e30;
  const int max_s = num_k / 8 > 0 ? num_k / 8 : 1;
  for (int s = 1; s <= max_s && s <= 64; ++s) {
    const int units = tiles * s;
    const int waves = (units + clusters - 1) / clusters;
    // cost ~ time: waves x per-unit k-work, plus a per-unit epilogue charge (~6 k-blocks worth)
    const double cost = (double)waves * ((double)num_k / s + 6.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_pair(const void* a, const void* b, int M, int N, int K, int splits, const CUtensorMap& tc_, const CUtensorMap& tc2_,
                const Epi& epi, cudaStream_t st) {
  CUtensorMap ta, tb;
  if (!A_MN) { if (int rc = make_tmap_2d(&ta, a, 2, (uint64_t)M, (uint64_t)K, (uint64_t)K * 2, kPM, 64)) return rc; }
  else       { if (int rc = make_tmap_2d(&ta, a, 2, (uint64_t)K, (uint64_t)M, (uint64_t)M * 2, 64, 64)) return rc; }
  if (!B_MN) { if (int rc = make_tmap_2d(&tb, b, 2, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, BN / 2, 64)) return rc; }
  else       { if (int rc = make_tmap_2d(&tb, b, 2, (uint64_t)K, (uint64_t)N, (uint64_t)N * 2, 64, 64)) return rc; }
  PairShape g;
  g.M = M; g.N = N; g.K = K;
  g.tiles_m = ceil_div(M, 2 * kPM);
  g.tiles_n = ceil_div(N, BN);
  g.num_k = ceil_div(K, kPK);
  g.splits = splits < 1 ? 1 : (splits > g.num_k ? g.num_k : splits);
  constexpr size_t smem = pair_smem_bytes<BN>();
  auto kern = gemm_pair_kernel<BN, A_MN, B_MN, Epi>;
  VRR_SMEM_ATTR_ONCE(kern, smem);
  const int clusters_max = sm_count() / 2;
  const long long units = (long long)g.tiles_m * g.tiles_n * g.splits;
  const int clusters = (int)(units < clusters_max ? units : clusters_max);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VRR_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc_, tc2_, g, epi));
  VRR_LAUNCHED();
  return VRR_OK;
}

template <bool A_MN, bool B_MN>
int launch_store(const void* a, const void* b, void* c, void* c2, const float* bias, int M, int N, int K, int c_f32,
                 int epilogue, int accumulate, float* colsum, cudaStream_t st) {
  StoreEpi epi;
  epi.colsum = colsum;
  epi.bias = (epilogue != VRR_EPI_NONE && epilogue != VRR_EPI_MUL) ? bias : nullptr;
  epi.out_f32 = c_f32;
  epi.gelu = epilogue == VRR_EPI_BIAS_GELU ? 1 : (epilogue == VRR_EPI_BIAS_GELU_GRAD ? 2 : (epilogue == VRR_EPI_BIAS_GELU_ACT ? 3 : 0));
  epi.mul = epilogue == VRR_EPI_MUL ? (const __nv_bfloat16*)c2 : nullptr;
  if (epilogue == VRR_EPI_MUL) epi.bias = nullptr;
  const int tiles = ceil_div(M, 2 * kPM) * ceil_div(N, N > 128 ? 256 : 128);
  int splits = 1;
  if (c_f32) splits = pick_splits(tiles, ceil_div(K, kPK), sm_count() / 2);
  epi.reduce = (c_f32 && (accumulate || splits > 1)) ? 1 : 0;
  if (c_f32 && !accumulate && splits > 1) VRR_CUDA(cudaMemsetAsync(c, 0, (size_t)M * N * sizeof(float), st));
  CUtensorMap tc_, tc2_;
  if (c_f32) {
    if (int rc = make_tmap_2d(&tc_, c, 4, (uint64_t)M, (uint64_t)N, (uint64_t)N * 4, 32, 32)) return rc;
    tc2_ = tc_;
  } else {
    if (int rc = make_tmap_2d(&tc_, c, 2, (uint64_t)M, (uint64_t)N, (uint64_t)N * 2, 32, 64)) return rc;
    tc2_ = tc_;
    if (epi.gelu == 1 || epi.gelu == 2)
      if (int rc = make_tmap_2d(&tc2_, c2, 2, (uint64_t)M, (uint64_t)N, (uint64_t)N * 2, 32, 64)) return rc;
  }
  if (N > 128) return launch_pair<256, A_MN, B_MN, StoreEpi>(a, b, M, N, K, splits, tc_, tc2_, epi, st);
  return launch_pair<128, A_MN, B_MN, StoreEpi>(a, b, M, N, K, splits, tc_, tc2_, epi, st);
}

std::atomic<int> g_gemm_variant{2};

}  // namespace

void gemm_tc_set_variant(int v) { g_gemm_variant.store(v == 1 ? 1 : 2); }
int gemm_tc_variant() { return g_gemm_variant.load(); }

// bf16 operands; TMA needs 16-byte row strides: the CONTIGUOUS dimension of every operand and of C must be a
// multiple of 8 elements (4 for an fp32 C); the other dimensions are free (ragged tiles are zero-filled / clipped).
bool gemm_bf16_tc_supported(int M, int N, int K, int trans_a, int trans_b, int c_dtype, int epilogue) {
  if (M < 1 || N < 1 || K < 1) return false;
  if ((trans_a ? M : K) % 8 != 0) return false;   // A stored [K][M] or [M][K]
  if ((trans_b ? K : N) % 8 != 0) return false;   // B stored [N][K] or [K][N]
  if (c_dtype == VRR_BF16 ? (N % 8 != 0) : (N % 4 != 0)) return false;
  if (c_dtype != VRR_F32 && c_dtype != VRR_BF16) return false;
  if (epilogue >= VRR_EPI_BIAS_GELU && c_dtype != VRR_BF16) return false;
  if (epilogue != VRR_EPI_NONE && N % 8 != 0) return false;  // vector loads of the bias / the multiplier
  if ((long long)M * N >= (1ll << 40)) return false;
  return true;
}

int gemm_bf16_tc(const void* a, const void* b, void* c, void* c2, const float* bias, int M, int N, int K, int trans_a,
                 int trans_b, int c_dtype, int epilogue, int accumulate, cudaStream_t st, float* colsum) {
  VRR_REQUIRE(colsum == nullptr || (c_dtype == VRR_BF16 && epilogue != VRR_EPI_BIAS_GELU && epilogue != VRR_EPI_BIAS_GELU_GRAD &&
                                    epilogue != VRR_EPI_BIAS_GELU_ACT),
              VRR_ERR_UNSUPPORTED, "gemm (tcgen05): column sums come with the bf16 plain / bias / multiply epilogues only");
  VRR_REQUIRE((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)c2) & 15) == 0, VRR_ERR_INVALID_ARG,
              "gemm (tcgen05): a / b / c must be 16-byte aligned");
  VRR_REQUIRE(epilogue == VRR_EPI_NONE || epilogue == VRR_EPI_MUL || (bias != nullptr && ((uintptr_t)bias & 15) == 0),
              VRR_ERR_INVALID_ARG, "gemm (tcgen05): the bias epilogues need a 16-byte aligned fp32 bias");
  VRR_REQUIRE(epilogue < VRR_EPI_BIAS_GELU || epilogue == VRR_EPI_BIAS_GELU_ACT || c2 != nullptr, VRR_ERR_INVALID_ARG,
              "gemm (tcgen05): the two-output GELU epilogues and MUL need c2");
  VRR_REQUIRE(!accumulate || c_dtype == VRR_F32, VRR_ERR_UNSUPPORTED, "gemm (tcgen05): accumulate needs an fp32 C");
  const int c_f32 = c_dtype == VRR_F32;
  const bool a_mn = trans_a != 0, b_mn = trans_b == 0;
  if (!a_mn && !b_mn) return launch_store<false, false>(a, b, c, c2, bias, M, N, K, c_f32, epilogue, accumulate, colsum, st);
  if (!a_mn && b_mn) return launch_store<false, true>(a, b, c, c2, bias, M, N, K, c_f32, epilogue, accumulate, colsum, st);
  if (a_mn && b_mn) return launch_store<true, true>(a, b, c, c2, bias, M, N, K, c_f32, epilogue, accumulate, colsum, st);
  return launch_store<true, false>(a, b, c, c2, bias, M, N, K, c_f32, epilogue, accumulate, colsum, st);
}

// ---- QKV projection + RoPE epilogue on the CTA-pair kernel -----------------------------------------------
int qkv_rope_fwd_tc2(const void* x, const void* w, const float* cos_tab, const float* sin_tab, const float* packed,
                     void* planes, int B, int N, int E, int H, int rope_mode, cudaStream_t st) {
  VRR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)planes & 15) == 0,
              VRR_ERR_INVALID_ARG, "qkv_rope_fwd (tcgen05): x / w / planes must be 16-byte aligned");
  if (rope_mode != VRR_ROPE_NONE)
    VRR_REQUIRE(((uintptr_t)cos_tab & 15) == 0 && ((uintptr_t)sin_tab & 15) == 0, VRR_ERR_INVALID_ARG,
                "qkv_rope_fwd (tcgen05): cos / sin must be 16-byte aligned");
  VRR_REQUIRE(((uintptr_t)packed & 15) == 0, VRR_ERR_INVALID_ARG, "qkv_rope_fwd (tcgen05): packed table must be 16-byte aligned");
  QkvRopeEpi2 epi{(__nv_bfloat16*)planes, cos_tab, sin_tab, reinterpret_cast<const float4*>(packed), B, N, E, H, rope_mode};
  CUtensorMap tplanes;  // planes[3 B H][N][64] bf16, box = [1][32 rows][64]
  if (int rc = make_tmap_3d_bf16(&tplanes, planes, 64, (uint64_t)N, (uint64_t)3 * B * H, 128, (uint64_t)N * 128, 64, 32)) return rc;
  // 64-column chunks are whole heads (3E = 192 H); a ragged last n-tile is clipped by the n0 < N test
  if (3 * E > 128) return launch_pair<256, false, false, QkvRopeEpi2>(x, w, B * N, 3 * E, E, 1, tplanes, tplanes, epi, st);
  return launch_pair<128, false, false, QkvRopeEpi2>(x, w, B * N, 3 * E, E, 1, tplanes, tplanes, epi, st);
}

int patch_embed_gemm_tc2(const void* unfolded, const void* weight, const void* bias, const void* pos, void* tokens, int M,
                         int Np, int K, int E, int tok_dtype, cudaStream_t st) {
  CUtensorMap dummy;
  memset(&dummy, 0, sizeof(dummy));
  if (tok_dtype == VRR_F32) {
    PatchEmbedEpi2<float> epi{(float*)tokens, (const __nv_bfloat16*)bias, (const float*)pos, Np, E};
    if (E > 128) return launch_pair<256, false, false, PatchEmbedEpi2<float>>(unfolded, weight, M, E, K, 1, dummy, dummy, epi, st);
    return launch_pair<128, false, false, PatchEmbedEpi2<float>>(unfolded, weight, M, E, K, 1, dummy, dummy, epi, st);
  }
  PatchEmbedEpi2<__nv_bfloat16> epi{(__nv_bfloat16*)tokens, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)pos, Np, E};
  if (E > 128)
    return launch_pair<256, false, false, PatchEmbedEpi2<__nv_bfloat16>>(unfolded, weight, M, E, K, 1, dummy, dummy, epi, st);
  return launch_pair<128, false, false, PatchEmbedEpi2<__nv_bfloat16>>(unfolded, weight, M, E, K, 1, dummy, dummy, epi, st);
}

}  // namespace vrr
