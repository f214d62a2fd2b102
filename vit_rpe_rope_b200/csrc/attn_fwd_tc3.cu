// Fused attention forward, variant 3 (default): PERSISTENT, one CTA per SM, two softmax warpgroups.
//
// Replaces models/vit.py:71-88 of the reference (S = q k^T * scale (+ bias) -> softmax -> . v -> merge heads).
// Same math and TMEM operand forms as variant 2 (attn_tc.cu); what changes is the schedule.  At 197 tokens
// variant 2 spent ~85 % of each CTA's life in fixed cost (launch, barrier init, TMEM allocation, the first
// TMA round trip, the epilogue) because a CTA lived for four key tiles only.  Here
//   * the grid is one CTA per SM; each CTA loops over work items (image, head, pair of 128-row query tiles);
//   * warp 8 (one lane) is the TMA producer: Q (double-buffered per item) and 64-key K / V tiles through an
//     8-stage ring, running up to two items ahead of the tensor core - the load latency of item k+1 hides
//     behind item k;
//   * warp 9 (one lane) issues every tcgen05.mma for BOTH query tiles, one S tile ahead of the softmax,
//     across item boundaries;
//   * warps 0-3 / 4-7 are two softmax warpgroups (thread = query row = TMEM lane), one per 128-row tile of
//     the item: K and V are staged once for both, and the tensor core works for one group while the other
//     is in its exponentials;
//   * TMEM (all 512 columns): per group  S/P[2] | O[2] - O is double-buffered across items so the
//     epilogue of item k (O / l -> bf16 -> HBM) overlaps the MMAs of item k+1;
//   * the 1-key tail tile (197 = 3*64 + 5, 577 = 9*64 + 1, ...) runs 16 columns wide (MMA N = 16) instead of 64.
// The relative-position table row of the item's head is staged by a TMA bulk copy issued by the producer
// warp one item ahead; the polynomial bias is a per-head distance LUT.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrr {

using namespace tc;

namespace {

constexpr int kDh = 64;
constexpr int kKT = 64;                    // keys per tile
constexpr int kStages = 8;                 // K ring and V ring depth
constexpr int kTileBytes = kKT * kDh * 2;  // 8 KB
constexpr int kQBufBytes = 256 * kDh * 2;  // 32 KB: both 128-row query tiles of one item
constexpr int kThreads = 320;
constexpr uint32_t kTmemCols = 512;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kNumBars = 4 * kStages + 4 + 8 + 2 + 8 + 2;
constexpr int kSmemMax = 232448;

struct Fwd3Params {
  __nv_bfloat16* out;
  float* lse;
  const float* bias_param;
  int B, H, N;
  float scale_log2;
  int bias_heads, bias_len, bias_grid;
  int lut_floats;   // floats of ONE group-private LUT
  int raw_floats;   // TABLE: floats of the TMA staging buffer
  float rescale_threshold;
  int ntiles, npairs, total_items;
  long long* dbg;  // optional clock64 stamps of CTA 0 (4 regions x 256 slots); NULL in production
};

// debug: stamp slot `idx` of region `region` (CTA 0 only)
__device__ __forceinline__ void dbg_stamp(const Fwd3Params& p, int region, int idx) {
  if (p.dbg != nullptr && blockIdx.x == 0 && idx < 256) p.dbg[region * 256 + idx] = clock64();
}

// aligned span of the table row of head h that a bulk copy may touch (rows of 2N-1 floats are not 16-byte
// aligned in general): [a0, a1) bytes from the table start; the row itself starts at row_b
__device__ __forceinline__ void table_span(const Fwd3Params& p, int h, size_t& row_b, size_t& a0, size_t& a1) {
  row_b = (size_t)h * p.bias_len * 4;
  const size_t row_e = row_b + (size_t)p.bias_len * 4;
  const size_t tot = (size_t)p.bias_heads * p.bias_len * 4;
  a0 = row_b & ~size_t(15);
  a1 = (row_e + 15) & ~size_t(15);
  if (a1 > (tot & ~size_t(15))) a1 = tot & ~size_t(15);
}

// Softmax of one key tile for one query row, in blocks of W (32 or 16) accumulator columns held in registers.
// pass 1: mask, scale + bias (bias modes), running tile max.
template <int BIAS, int W>
__device__ __forceinline__ void tile_pass1(const Fwd3Params& p, uint32_t (&s)[W], int nvalid, int i, int j0,
                                           const float* lut, const uint16_t* key_yx, int yi, int xi, float (&tm)[4]) {
#pragma unroll
  for (int jl = 0; jl < W; ++jl) {
    if (jl >= nvalid) s[jl] = 0xff800000u;  // -inf: key past the end of the sequence
    if (BIAS == VRR_BIAS_NONE) {
      tm[jl & 3] = fmaxf(tm[jl & 3], __uint_as_float(s[jl]));
    } else {
      const int j = j0 + jl;
      float bsv;
      if (BIAS == VRR_BIAS_TABLE) {
        bsv = lut[min(max(i - j + p.N - 1, 0), 2 * p.N - 2)];
      } else {
        const int yx = key_yx[min(j, p.N - 1)];
        const int dist = abs(yi - (yx >> 8)) + abs(xi - (yx & 255));
        bsv = (i == 0 || j == 0) ? 0.f : lut[dist];
      }
      const float v = fmaf(__uint_as_float(s[jl]), p.scale_log2, bsv);
      s[jl] = __float_as_uint(v);
      tm[jl & 3] = fmaxf(tm[jl & 3], v);
    }
  }
}
// pass 2: exponentials against the reference max, row sum, P (bf16 pairs) back over the S buffer.
template <int BIAS, int W>
__device__ __forceinline__ void tile_pass2(const Fwd3Params& p, const uint32_t (&s)[W], uint32_t t_dst, float neg_m,
                                           float (&ls)[2]) {
#pragma unroll
  for (int c = 0; c < W / 16; ++c) {
    uint32_t packed[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a0 = __uint_as_float(s[c * 16 + 2 * e]), a1 = __uint_as_float(s[c * 16 + 2 * e + 1]);
      const float p0 = BIAS == VRR_BIAS_NONE ? ex2(fmaf(a0, p.scale_log2, neg_m)) : ex2(a0 + neg_m);
      const float p1 = BIAS == VRR_BIAS_NONE ? ex2(fmaf(a1, p.scale_log2, neg_m)) : ex2(a1 + neg_m);
      ls[0] += p0;
      ls[1] += p1;
      packed[e] = pack_bf16(p0, p1);
    }
    tmem_st8(t_dst + c * 8, packed);
  }
}

// One key tile of COLS (64 or 16) columns for one query row: logits from TMEM -> bias -> lazy running max ->
// exponentials -> P (bf16) back over the S buffer.
template <int BIAS, int COLS>
__device__ __forceinline__ void softmax_tile(const Fwd3Params& p, uint32_t t_s, uint32_t t_o, uint64_t* bar_pv,
                                             uint32_t pv_parity, bool first_tile, int nvalid, int i, int j0,
                                             const float* lut, const uint16_t* key_yx, int yi, int xi, float& m_ref,
                                             float& l_run) {
  constexpr int W = COLS == 64 ? 32 : 16;
  uint32_t s0[W], s1[W];
  if (COLS == 64) {
    tmem_ld32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&s0));
    tmem_ld32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s1));
  } else {
    tmem_ld16(t_s, *reinterpret_cast<uint32_t(*)[16]>(&s0));
  }
  tmem_wait_ld();
  float tm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  tile_pass1<BIAS, W>(p, s0, nvalid, i, j0, lut, key_yx, yi, xi, tm);
  if (COLS == 64) tile_pass1<BIAS, W>(p, s1, nvalid - 32, i, j0 + 32, lut, key_yx, yi, xi, tm);
  float tmax = fmaxf(fmaxf(tm[0], tm[1]), fmaxf(tm[2], tm[3]));
  if (BIAS == VRR_BIAS_NONE) tmax *= p.scale_log2;
  // reference-max update: warp-uniform decision (tcgen05.ld / st are warp-collective)
  const bool jump = tmax > m_ref + p.rescale_threshold;
  if (__any_sync(0xffffffffu, jump)) {
    const float m_new = fmaxf(m_ref, tmax);
    if (!first_tile) {
      const float f = ex2(m_ref - m_new);  // 1 for rows whose reference did not move
      l_run *= f;
      // PV of the previous tile retired.  Safe parity wait: this tile's S was committed after the PV two
      // tiles back, so the barrier is at most one phase behind.
      mbar_wait(bar_pv, pv_parity);
      tc_fence_after();
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {  // rare path: 16 columns at a time keeps the register peak of the tile low
        uint32_t o[16];
        tmem_ld16(t_o + q * 16, o);
        tmem_wait_ld();
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * f);
        tmem_st16(t_o + q * 16, o);
      }
    }
    m_ref = m_new;
  }
  float ls[2] = {0.f, 0.f};
  tile_pass2<BIAS, W>(p, s0, t_s, -m_ref, ls);
  if (COLS == 64) tile_pass2<BIAS, W>(p, s1, t_s + 16, -m_ref, ls);
  l_run += ls[0] + ls[1];
  tmem_wait_st();
}

template <int BIAS>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Fwd3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // [2 items][256 rows][128 B]
  uint8_t* sK = sQ + 2 * kQBufBytes;                // [kStages][8 KB]
  uint8_t* sV = sK + kStages * kTileBytes;          // [kStages][8 KB]
  float* lut_raw = reinterpret_cast<float*>(sV + kStages * kTileBytes);  // TABLE: TMA staging of one table row
  float* lut_wg = lut_raw + p.raw_floats;                                // [2 groups][lut_floats]
  uint16_t* key_yx = reinterpret_cast<uint16_t*>(lut_wg + 2 * p.lut_floats);
  const int key_yx_bytes = (BIAS == VRR_BIAS_POLY) ? ((p.N * 2 + 15) & ~15) : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(key_yx) + key_yx_bytes);
  uint64_t* bar_kfull = bars;                       // [kStages] TMA -> issuer
  uint64_t* bar_kempty = bars + kStages;            // [kStages] S MMAs of the tile retired -> producer
  uint64_t* bar_vfull = bars + 2 * kStages;
  uint64_t* bar_vempty = bars + 3 * kStages;        // PV MMAs of the tile retired -> producer
  uint64_t* bar_qfull = bars + 4 * kStages;         // [2]
  uint64_t* bar_qempty = bar_qfull + 2;             // [2] last S MMA of the item retired
  uint64_t* bar_sfull = bar_qempty + 2;             // [2 groups][2 buffers] S ready (issuer commit)
  uint64_t* bar_pfull = bar_sfull + 4;              // [2][2] P stored (128 arrivals)
  uint64_t* bar_pv = bar_pfull + 4;                 // [2] one phase per PV of the group (lazy rescale)
  uint64_t* bar_ofull = bar_pv + 2;                 // [2][2] last PV of the item retired
  uint64_t* bar_oempty = bar_ofull + 4;             // [2][2] epilogue has read O (128 arrivals)
  uint64_t* bar_lutfull = bar_oempty + 4;           // table row landed
  uint64_t* bar_lutempty = bar_lutfull + 1;         // 256 arrivals: staging buffer copied out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_lutempty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H;
  const int ntiles = p.ntiles, npairs = p.npairs, total = p.total_items;
  const int BHN = p.B * H * N;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_kfull[s], 1);
      mbar_init(&bar_kempty[s], 1);
      mbar_init(&bar_vfull[s], 1);
      mbar_init(&bar_vempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_qfull[s], 1);
      mbar_init(&bar_qempty[s], 1);
      mbar_init(&bar_pv[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&bar_sfull[s], 1);
      mbar_init(&bar_pfull[s], 128);
      mbar_init(&bar_ofull[s], 1);
      mbar_init(&bar_oempty[s], 128);
    }
    mbar_init(bar_lutfull, 1);
    mbar_init(bar_lutempty, 256);
    fence_mbar_init();
  }
  if (BIAS == VRR_BIAS_POLY) {
    for (int t = tid; t < N; t += kThreads) {
      const int pt = t > 0 ? t - 1 : 0;
      key_yx[t] = (uint16_t)(((pt % p.bias_grid) << 8) | (pt / p.bias_grid));
    }
  }
  __syncwarp();
  if (warp == 9) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The producer and issuer warps run CONVERGED (all 32 lanes execute the loops and the waits) and issue under
  // elect_one(): ptxas then keeps descriptors / addresses in uniform registers and emits back-to-back UTCHMMA /
  // UTMALDG.  Under `if (lane == 0)` it wraps every tcgen05 / TMA instruction in an elect-and-branch loop
  // (~100 cycles per instruction, measured with the clock64 stamps below: the issuer was the bottleneck).
  if (warp == 8) {
    // ============================================ TMA producer ============================================
    if (elect_one()) tma_prefetch_desc(&tmap);
    __syncwarp();
    uint32_t f = 0;
    int k = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++k) {
      const int bh = item / npairs, rp = item - bh * npairs;
      const int qb = k & 1;
      mbar_wait(&bar_qempty[qb], (uint32_t)(((k >> 1) & 1) ^ 1));
      const int nbox = (min(256, N - rp * 256) + 63) >> 6;
      if (elect_one()) {
        mbar_expect_tx(&bar_qfull[qb], (uint32_t)nbox * kTileBytes);
        for (int j = 0; j < nbox; ++j)
          tma_load_2d(sQ + qb * kQBufBytes + j * kTileBytes, &tmap, &bar_qfull[qb], 0, bh * N + rp * 256 + j * 64);
      }
      __syncwarp();
      if (BIAS == VRR_BIAS_TABLE) {
        mbar_wait(bar_lutempty, (uint32_t)((k & 1) ^ 1));
        size_t row_b, a0, a1;
        table_span(p, bh % H, row_b, a0, a1);
        if (elect_one()) {
          if (a1 > a0) {
            mbar_expect_tx(bar_lutfull, (uint32_t)(a1 - a0));
            bulk_load_1d(lut_raw, reinterpret_cast<const uint8_t*>(p.bias_param) + a0, (uint32_t)(a1 - a0), bar_lutfull);
          } else {
            mbar_arrive(bar_lutfull);
          }
        }
        __syncwarp();
      }
      for (int t = 0; t < ntiles; ++t, ++f) {
        const int st = f % kStages;
        const uint32_t ph = (f / kStages) & 1;
        mbar_wait(&bar_kempty[st], ph ^ 1);
        if (lane == 0) dbg_stamp(p, 3, (int)f * 2);
        if (elect_one()) {
          mbar_expect_tx(&bar_kfull[st], kTileBytes);
          tma_load_2d(sK + st * kTileBytes, &tmap, &bar_kfull[st], 0, BHN + bh * N + t * kKT);
        }
        __syncwarp();
        mbar_wait(&bar_vempty[st], ph ^ 1);
        if (lane == 0) dbg_stamp(p, 3, (int)f * 2 + 1);
        if (elect_one()) {
          mbar_expect_tx(&bar_vfull[st], kTileBytes);
          tma_load_2d(sV + st * kTileBytes, &tmap, &bar_vfull[st], 0, 2 * BHN + bh * N + t * kKT);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ============================================ MMA issuer ==============================================
    const int n_local = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int total_tiles = n_local * ntiles;
    uint32_t gs0 = 0, gs1 = 0, gp0 = 0, gp1 = 0, kw0 = 0, kw1 = 0;
    constexpr uint32_t idesc_o = idesc_bf16(128, kDh, 0, 1);
    const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV);
    // S(k, t) for both row tiles of the item; (ks, ts, rps) walk one tile ahead of (k, t, rp)
    int ks = 0, ts = 0, rps = (int)(blockIdx.x % npairs);
    auto issue_s = [&](int f) {
      if (ts == 0) mbar_wait(&bar_qfull[ks & 1], (uint32_t)((ks >> 1) & 1));
      const int st = f % kStages;
      mbar_wait(&bar_kfull[st], (uint32_t)((f / kStages) & 1));
      tc_fence_after();
      const int ncols = (min(kKT, N - ts * kKT) + 15) & ~15;
      const uint32_t idesc_s = idesc_bf16(128, ncols, 0, 0);
      const bool act1 = rps * 256 + 128 < N;
      if (elect_one()) {
        const uint64_t dk = smem_desc_sw128(k_base + st * kTileBytes);
        const uint64_t dq0 = smem_desc_sw128(q_base + (ks & 1) * kQBufBytes);
        const uint32_t d0 = tmem_base + (gs0 & 1) * kKT;
#pragma unroll
        for (int kk = 0; kk < kDh / 16; ++kk) mma_ss(d0, dq0 + 2 * kk, dk + 2 * kk, idesc_s, kk > 0);
        mma_commit(&bar_sfull[gs0 & 1]);
        if (act1) {
          const uint64_t dq1 = smem_desc_sw128(q_base + (ks & 1) * kQBufBytes + 128 * kDh * 2);
          const uint32_t d1 = tmem_base + 256 + (gs1 & 1) * kKT;
#pragma unroll
          for (int kk = 0; kk < kDh / 16; ++kk) mma_ss(d1, dq1 + 2 * kk, dk + 2 * kk, idesc_s, kk > 0);
          mma_commit(&bar_sfull[2 + (gs1 & 1)]);
        }
        mma_commit(&bar_kempty[st]);
        if (ts == ntiles - 1) mma_commit(&bar_qempty[ks & 1]);
      }
      __syncwarp();
      ++gs0;
      if (act1) ++gs1;
      if (++ts == ntiles) {
        ts = 0;
        ++ks;
        rps = (int)((blockIdx.x + (unsigned)ks * gridDim.x) % npairs);
      }
    };
    if (total_tiles > 0) issue_s(0);
    int k = 0, t = 0, rp = (int)(blockIdx.x % npairs);
    for (int f = 0; f < total_tiles; ++f) {
      if (lane == 0) dbg_stamp(p, 2, f * 8 + 0);
      if (f + 1 < total_tiles) issue_s(f + 1);  // one S tile ahead of the softmax, across item boundaries
      if (lane == 0) dbg_stamp(p, 2, f * 8 + 1);
      const int st = f % kStages;
      const bool act1 = rp * 256 + 128 < N;
      const bool last = t == ntiles - 1;
      const int ksteps = (min(kKT, N - t * kKT) + 15) >> 4;
      mbar_wait(&bar_vfull[st], (uint32_t)((f / kStages) & 1));
      if (lane == 0) dbg_stamp(p, 2, f * 8 + 2);
      // ---- row tile 0
      {
        const uint32_t ob = kw0 & 1, sb = gp0 & 1;
        if (t == 0) mbar_wait(&bar_oempty[ob], ((kw0 >> 1) & 1) ^ 1);  // epilogue two items back has read O
        mbar_wait(&bar_pfull[sb], (gp0 >> 1) & 1);
        if (lane == 0) dbg_stamp(p, 2, f * 8 + 3);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dv = smem_desc_sw128(v_base + st * kTileBytes);
          const uint32_t a = tmem_base + sb * kKT, d = tmem_base + 128 + ob * kDh;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (kk < ksteps) mma_ts(d, a + kk * 8, dv + 128 * kk, idesc_o, (t | kk) != 0);
          mma_commit(&bar_pv[0]);
          if (last) mma_commit(&bar_ofull[ob]);
          if (!act1) mma_commit(&bar_vempty[st]);
        }
        __syncwarp();
        if (lane == 0) dbg_stamp(p, 2, f * 8 + 4);
        ++gp0;
        if (last) ++kw0;
      }
      // ---- row tile 1
      if (act1) {
        const uint32_t ob = kw1 & 1, sb = gp1 & 1;
        if (t == 0) mbar_wait(&bar_oempty[2 + ob], ((kw1 >> 1) & 1) ^ 1);
        mbar_wait(&bar_pfull[2 + sb], (gp1 >> 1) & 1);
        if (lane == 0) dbg_stamp(p, 2, f * 8 + 5);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dv = smem_desc_sw128(v_base + st * kTileBytes);
          const uint32_t a = tmem_base + 256 + sb * kKT, d = tmem_base + 256 + 128 + ob * kDh;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (kk < ksteps) mma_ts(d, a + kk * 8, dv + 128 * kk, idesc_o, (t | kk) != 0);
          mma_commit(&bar_pv[1]);
          if (last) mma_commit(&bar_ofull[2 + ob]);
          mma_commit(&bar_vempty[st]);
        }
        __syncwarp();
        if (lane == 0) dbg_stamp(p, 2, f * 8 + 6);
        ++gp1;
        if (last) ++kw1;
      }
      if (++t == ntiles) {
        t = 0;
        ++k;
        rp = (int)((blockIdx.x + (unsigned)k * gridDim.x) % npairs);
      }
    }
  } else {
    // ============================================ softmax groups ==========================================
    const int w = warp >> 2, wq = warp & 3, r = tid & 127;
    const uint32_t tmem_wg = tmem_base + (uint32_t)(w * 256) + ((uint32_t)(wq * 32) << 16);
    float* lut = lut_wg + w * p.lut_floats;
    uint32_t gw = 0, kw = 0;
    bool lut_valid = false;
    int k = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++k) {
      const int bh = item / npairs, rp = item - bh * npairs;
      const int b = bh / H, h = bh - b * H;
      const int m0 = rp * 256 + w * 128;
      const bool active = m0 < N;
      const int i = m0 + r;
      int yi = 0, xi = 0;
      if (BIAS == VRR_BIAS_TABLE) {
        mbar_wait(bar_lutfull, (uint32_t)(k & 1));
        if (active) {
          size_t row_b, a0, a1;
          table_span(p, h, row_b, a0, a1);
          const int have = a1 > row_b ? (int)((a1 - row_b) / 4) : 0;  // elements the bulk copy delivered
          const float* src = lut_raw + (row_b - a0) / 4;
          const float* grow = p.bias_param + (size_t)h * p.bias_len;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");  // every row of the group left the previous item
          for (int t = r; t < p.bias_len; t += 128) lut[t] = (t < have ? src[t] : __ldg(grow + t)) * kLog2e;
        }
        mbar_arrive(bar_lutempty);
        if (active) asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
      } else if (BIAS == VRR_BIAS_POLY) {
        if (active && (!lut_valid || p.bias_heads > 1)) {
          const float* c = p.bias_param + (size_t)(p.bias_heads == 1 ? 0 : h) * p.bias_len;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
          for (int d = r; d < 2 * p.bias_grid - 1; d += 128) {
            float x = (float)d, pw = 1.f, acc = 0.f;
            for (int q = 0; q < p.bias_len; ++q) {
              acc = fmaf(pw, c[q], acc);
              pw *= x;
            }
            lut[d] = acc * kLog2e;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
          lut_valid = true;
        }
        const int pi = i > 0 ? i - 1 : 0;
        yi = pi % p.bias_grid;
        xi = pi / p.bias_grid;
      }
      if (!active) continue;
      const bool warp_active = (m0 + wq * 32) < N;

      float m_ref = -INFINITY, l_run = 0.f;
      const uint32_t ob = kw & 1;
      const uint32_t t_o = tmem_wg + 128 + ob * kDh;
      for (int t = 0; t < ntiles; ++t, ++gw) {
        const uint32_t sb = gw & 1;
        const int nvalid = min(kKT, N - t * kKT);
        const bool dbg_on = p.dbg != nullptr && r == 0;
        if (dbg_on) dbg_stamp(p, w, (int)gw * 4 + 0);
        mbar_wait(&bar_sfull[w * 2 + sb], (gw >> 1) & 1);
        tc_fence_after();
        if (dbg_on) dbg_stamp(p, w, (int)gw * 4 + 1);
        if (warp_active) {
          const uint32_t t_s = tmem_wg + sb * kKT;
          if (nvalid <= 16)
            softmax_tile<BIAS, 16>(p, t_s, t_o, &bar_pv[w], (gw - 1) & 1, t == 0, nvalid, i, t * kKT, lut, key_yx, yi, xi,
                                   m_ref, l_run);
          else
            softmax_tile<BIAS, 64>(p, t_s, t_o, &bar_pv[w], (gw - 1) & 1, t == 0, nvalid, i, t * kKT, lut, key_yx, yi, xi,
                                   m_ref, l_run);
        }
        if (dbg_on) dbg_stamp(p, w, (int)gw * 4 + 2);
        tc_fence_before();
        mbar_arrive(&bar_pfull[w * 2 + sb]);
        if (dbg_on) dbg_stamp(p, w, (int)gw * 4 + 3);
      }

      // ---- epilogue: O / l -> bf16 -> out[b][i][h*64 ..], lse ----------------------------------------
      mbar_wait(&bar_ofull[w * 2 + ob], (kw >> 1) & 1);
      tc_fence_after();
      uint32_t olo[32], ohi[32];
      if (warp_active) {
        tmem_ld32(t_o, olo);
        tmem_ld32(t_o + 32, ohi);
        tmem_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&bar_oempty[w * 2 + ob]);
      ++kw;
      if (warp_active && i < N) {
        const float inv = 1.f / l_run;
        __nv_bfloat16* dst = p.out + ((size_t)b * N + i) * (size_t)(H * kDh) + h * kDh;
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          uint4 o4;
          o4.x = pack_bf16(__uint_as_float(olo[v8 * 8 + 0]) * inv, __uint_as_float(olo[v8 * 8 + 1]) * inv);
          o4.y = pack_bf16(__uint_as_float(olo[v8 * 8 + 2]) * inv, __uint_as_float(olo[v8 * 8 + 3]) * inv);
          o4.z = pack_bf16(__uint_as_float(olo[v8 * 8 + 4]) * inv, __uint_as_float(olo[v8 * 8 + 5]) * inv);
          o4.w = pack_bf16(__uint_as_float(olo[v8 * 8 + 6]) * inv, __uint_as_float(olo[v8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + v8 * 8) = o4;
        }
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          uint4 o4;
          o4.x = pack_bf16(__uint_as_float(ohi[v8 * 8 + 0]) * inv, __uint_as_float(ohi[v8 * 8 + 1]) * inv);
          o4.y = pack_bf16(__uint_as_float(ohi[v8 * 8 + 2]) * inv, __uint_as_float(ohi[v8 * 8 + 3]) * inv);
          o4.z = pack_bf16(__uint_as_float(ohi[v8 * 8 + 4]) * inv, __uint_as_float(ohi[v8 * 8 + 5]) * inv);
          o4.w = pack_bf16(__uint_as_float(ohi[v8 * 8 + 6]) * inv, __uint_as_float(ohi[v8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 32 + v8 * 8) = o4;
        }
        p.lse[(size_t)bh * N + i] = (m_ref + log2f(l_run)) * kLn2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, kTmemCols);
}

size_t fwd3_smem_bytes(int N, const vrr_bias_desc* bias, int* lut_floats, int* raw_floats) {
  int lf = 0, rf = 0;
  const int mode = bias ? bias->mode : VRR_BIAS_NONE;
  if (mode == VRR_BIAS_TABLE) {
    lf = (2 * N - 1 + 3) & ~3;
    rf = (2 * N - 1 + 8 + 3) & ~3;  // + slack for the aligned span
  } else if (mode == VRR_BIAS_POLY) {
    lf = (2 * bias->grid - 1 + 3) & ~3;
  }
  if (lut_floats) *lut_floats = lf;
  if (raw_floats) *raw_floats = rf;
  const size_t yx = mode == VRR_BIAS_POLY ? (size_t)((N * 2 + 15) & ~15) : 0;
  return 1024 + (size_t)2 * kQBufBytes + (size_t)2 * kStages * kTileBytes + (size_t)(rf + 2 * lf) * 4 + yx +
         (size_t)kNumBars * 8 + 16;
}

std::atomic<int> g_fwd3_thresh_x100{800};
std::atomic<long long*> g_fwd3_dbg{nullptr};

template <int MODE>
int fwd3_launch(const CUtensorMap& tmap, const Fwd3Params& p, int grid, size_t smem, cudaStream_t st) {
  VRR_SMEM_ATTR_ONCE(attn_fwd_tc3_kernel<MODE>, kSmemMax);
  attn_fwd_tc3_kernel<MODE><<<grid, kThreads, smem, st>>>(tmap, p);
  VRR_LAUNCHED();
  return VRR_OK;
}

}  // namespace

bool attn_fwd_tc3_supported(int B, int H, int N, int Dh, const vrr_bias_desc* bias) {
  if (Dh != kDh || N < 1) return false;
  if ((long long)3 * B * H * N >= (1ll << 31)) return false;
  if ((long long)B * H * ((N + 255) / 256) >= (1ll << 30)) return false;
  if (bias && bias->mode == VRR_BIAS_POLY && bias->grid > 255) return false;
  return fwd3_smem_bytes(N, bias, nullptr, nullptr) <= (size_t)kSmemMax;
}

void attn_fwd_tc3_set_threshold_x100(int v) { g_fwd3_thresh_x100.store(v); }
void attn_fwd_tc3_set_debug(long long* buf) { g_fwd3_dbg.store(buf); }

int attn_fwd_tc3(const void* planes, const vrr_bias_desc* bias, void* out, float* lse, int B, int H, int N, int Dh,
                 float scale, cudaStream_t st) {
  (void)Dh;
  VRR_REQUIRE(((uintptr_t)planes & 15) == 0 && ((uintptr_t)out & 15) == 0, VRR_ERR_INVALID_ARG,
              "attn_fwd (tcgen05): planes/out must be 16-byte aligned");
  CUtensorMap tmap;
  if (int rc = make_tmap_bf16(&tmap, planes, (uint64_t)3 * B * H * N, kDh, kDh * 2, kKT)) return rc;
  Fwd3Params p;
  p.out = (__nv_bfloat16*)out;
  p.lse = lse;
  p.B = B; p.H = H; p.N = N;
  p.scale_log2 = scale * kLog2e;
  p.rescale_threshold = g_fwd3_thresh_x100.load() * 0.01f;
  p.dbg = g_fwd3_dbg.load();
  const int mode = bias ? bias->mode : VRR_BIAS_NONE;
  p.bias_param = mode != VRR_BIAS_NONE ? bias->param : nullptr;
  p.bias_heads = mode != VRR_BIAS_NONE ? bias->heads : 0;
  p.bias_len = mode != VRR_BIAS_NONE ? bias->len : 0;
  p.bias_grid = mode != VRR_BIAS_NONE ? bias->grid : 0;
  if (mode == VRR_BIAS_TABLE)
    VRR_REQUIRE(((uintptr_t)bias->param & 15) == 0, VRR_ERR_INVALID_ARG, "attn_fwd (tcgen05): bias table must be 16-byte aligned");
  p.ntiles = ceil_div(N, kKT);
  p.npairs = ceil_div(N, 256);
  p.total_items = B * H * p.npairs;
  const size_t smem = fwd3_smem_bytes(N, bias, &p.lut_floats, &p.raw_floats);
  const int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  if (mode == VRR_BIAS_TABLE) return fwd3_launch<VRR_BIAS_TABLE>(tmap, p, grid, smem, st);
  if (mode == VRR_BIAS_POLY) return fwd3_launch<VRR_BIAS_POLY>(tmap, p, grid, smem, st);
  return fwd3_launch<VRR_BIAS_NONE>(tmap, p, grid, smem, st);
}

}  // namespace vrr
