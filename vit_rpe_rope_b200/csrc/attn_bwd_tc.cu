// Fused attention backward on tcgen05 tensor cores (TMEM accumulators, TMA operand staging), bf16,
// head dim 64.  Autograd of models/vit.py:71-88 of the reference without ever
// materialising S, P, dP or dS in HBM:
//     delta = rowsum(dO * O)          P  = exp(q k^T * scale + bias - lse)
//     dP    = dO v^T                  dS = P * (dP - delta)
//     dV    = P^T dO                  dK = scale * dS^T q            dQ = scale * dS k
//     dTable[h][i-j+N-1] += dS        dcoef[k] += dS * dist(i,j)^k
//
// Two kernels, both shaped like the forward kernel (one thread = one row = one TMEM lane), so every
// MMA uses the operand forms the forward kernel proves: A,B from 128-byte-swizzled shared memory
// (K-major), or A from TMEM with B = [row][64] shared memory consumed MN-major.
//
//  (1) dQ kernel  - CTA = 128 query rows; K/V stream through a ring in 64-key tiles.
//        S = Q K^T, dP = dO V^T (SS MMAs, TMEM cols [0,64) / [64,128));  thread: P, dS (fp32),
//        bias-gradient reduction, dS -> bf16 -> TMEM;  dQ += dS K (TS MMA, accumulates in TMEM
//        cols [128,192) across all key tiles).  Also writes delta for kernel (2).
//        Relative-table gradient: a systolic warp-shuffle chain carries diagonal partial sums from
//        lane to lane, so a warp issues one shared-memory atomic per column instead of 32.
//        Polynomial gradient: per-thread power sums  sum_j dS*d^k  in registers.
//  (2) dK/dV kernel - CTA = 128 key rows; Q/dO stream through the ring in 64-query tiles.
//        S^T = K Q^T, dP^T = V dO^T;  thread: P^T, dS^T -> bf16 -> TMEM;
//        dV += P^T dO, dK += dS^T Q (TS MMAs, accumulate in TMEM cols [128,192) / [192,256)).
// 256 TMEM columns and < 100 KB shared memory per CTA: two CTAs per SM.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrr {

using namespace tc;

namespace {

constexpr int kRowsCta = 128;  // rows (lanes) per CTA
constexpr int kTile = 64;      // streamed rows per tile
constexpr int kRing = 3;
constexpr int kDh = 64;
constexpr int kTile64Bytes = kTile * kDh * 2;      // 8 KB
constexpr int kTile128Bytes = kRowsCta * kDh * 2;  // 16 KB
constexpr uint32_t kTmemCols = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct BwdParams {
  const __nv_bfloat16 *out, *d_out;  // [B][N][E]
  const float* lse;                  // [B][H][N]
  float* delta;                      // [B][H][N] workspace
  __nv_bfloat16* d_planes;           // [3][B][H][N][64]
  const float* bias_param;
  float* d_bias_param;               // TABLE: [H][2N-1]; POLY: [heads][len]; fp32, pre-zeroed
  int B, H, N;
  float scale, scale_log2;
  int bias_heads, bias_len, bias_grid;
  int lut_floats;
};

// Per-head bias LUT pre-multiplied by log2(e) (+ per-token packed coordinates for POLY).
template <int BIAS>
__device__ __forceinline__ void fill_lut(float* lut, uint16_t* key_yx, const BwdParams& p, int h) {
  if (BIAS == VRR_BIAS_TABLE) {
    const float* row = p.bias_param + (size_t)h * p.bias_len;
    for (int t = threadIdx.x; t < p.bias_len; t += blockDim.x) lut[t] = row[t] * kLog2e;
  } else if (BIAS == VRR_BIAS_POLY) {
    const float* c = p.bias_param + (size_t)(p.bias_heads == 1 ? 0 : h) * p.bias_len;
    for (int d = threadIdx.x; d < 2 * p.bias_grid - 1; d += blockDim.x) {
      float x = (float)d, pw = 1.f, acc = 0.f;
      for (int k = 0; k < p.bias_len; ++k) {
        acc = fmaf(pw, c[k], acc);
        pw *= x;
      }
      lut[d] = acc * kLog2e;
    }
    for (int t = threadIdx.x; t < p.N; t += blockDim.x) {
      const int pt = t > 0 ? t - 1 : 0;
      key_yx[t] = (uint16_t)(((pt % p.bias_grid) << 8) | (pt / p.bias_grid));
    }
  }
}

__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const uint32_t (&lo)[32], const uint32_t (&hi)[32],
                                               float mul) {
#pragma unroll
  for (int v8 = 0; v8 < 8; ++v8) {
    const uint32_t* src = v8 < 4 ? &lo[v8 * 8] : &hi[(v8 - 4) * 8];
    uint4 w;
    w.x = pack_bf16(__uint_as_float(src[0]) * mul, __uint_as_float(src[1]) * mul);
    w.y = pack_bf16(__uint_as_float(src[2]) * mul, __uint_as_float(src[3]) * mul);
    w.z = pack_bf16(__uint_as_float(src[4]) * mul, __uint_as_float(src[5]) * mul);
    w.w = pack_bf16(__uint_as_float(src[6]) * mul, __uint_as_float(src[7]) * mul);
    *reinterpret_cast<uint4*>(dst + v8 * 8) = w;
  }
}

// ============================================================================================ dQ
template <int BIAS>
__global__ void __launch_bounds__(128, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmap_pl, const __grid_constant__ CUtensorMap tmap_do,
                      const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                           // 128 x 64
  uint8_t* sG = sQ + kTile128Bytes;             // dO rows of this CTA, 128 x 64
  uint8_t* sRing = sG + kTile128Bytes;          // [kRing][K tile 8 KB | V tile 8 KB]
  float* lut = reinterpret_cast<float*>(sRing + kRing * 2 * kTile64Bytes);
  float* hist = lut + p.lut_floats;             // TABLE only: 2N-1 floats
  const int hist_floats = (BIAS == VRR_BIAS_TABLE) ? ((2 * p.N - 1 + 3) & ~3) : 0;
  uint16_t* key_yx = reinterpret_cast<uint16_t*>(hist + hist_floats);
  const int key_yx_bytes = (BIAS == VRR_BIAS_POLY) ? ((p.N * 2 + 15) & ~15) : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(key_yx) + key_yx_bytes);
  uint64_t* bar_full = bars;  // [kRing]
  uint64_t* bar_q = bars + kRing;
  uint64_t* bar_s = bars + kRing + 1;
  uint64_t* bar_o = bars + kRing + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kRing + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, H = p.H, E = H * kDh;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int m0 = blockIdx.x * kRowsCta;
  const int i = m0 + tid;
  const int ic = min(i, N - 1);
  const bool live = i < N;
  const bool warp_active = (m0 + warp * 32) < N;
  const int ntiles = (N + kTile - 1) / kTile;
  const int BHN = p.B * H * N;

  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) mbar_init(&bar_full[s], 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  fill_lut<BIAS>(lut, key_yx, p, h);
  if (BIAS == VRR_BIAS_TABLE)
    for (int t = tid; t < 2 * N - 1; t += 128) hist[t] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);

  if (tid == 0) {
    tma_prefetch_desc(&tmap_pl);
    tma_prefetch_desc(&tmap_do);
    mbar_expect_tx(bar_q, 2 * kTile128Bytes);
    tma_load_2d(sQ, &tmap_pl, bar_q, 0, bh * N + m0);
    tma_load_2d(sQ + kTile64Bytes, &tmap_pl, bar_q, 0, bh * N + m0 + kTile);
    tma_load_2d(sG, &tmap_do, bar_q, h * kDh, b * N + m0);
    tma_load_2d(sG + kTile64Bytes, &tmap_do, bar_q, h * kDh, b * N + m0 + kTile);
    for (int s = 0; s < kRing && s < ntiles; ++s) {
      uint8_t* st = sRing + s * 2 * kTile64Bytes;
      mbar_expect_tx(&bar_full[s], 2 * kTile64Bytes);
      tma_load_2d(st, &tmap_pl, &bar_full[s], 0, BHN + bh * N + s * kTile);
      tma_load_2d(st + kTile64Bytes, &tmap_pl, &bar_full[s], 0, 2 * BHN + bh * N + s * kTile);
    }
  }
  __syncwarp();

  // per-row statistics: lse (log2 domain) and delta = sum_d dO * O (also stored for the dK/dV kernel)
  float lse2 = 0.f, delta = 0.f;
  {
    const size_t off = ((size_t)b * N + ic) * E + h * kDh;
    const uint4* o4 = reinterpret_cast<const uint4*>(p.out + off);
    const uint4* g4 = reinterpret_cast<const uint4*>(p.d_out + off);
#pragma unroll
    for (int v8 = 0; v8 < 8; ++v8) {
      const uint4 ov = __ldg(o4 + v8), gv = __ldg(g4 + v8);
      const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[e]));
        const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]));
        delta = fmaf(of.x, gf.x, delta);
        delta = fmaf(of.y, gf.y, delta);
      }
    }
    lse2 = p.lse[(size_t)bh * N + ic] * kLog2e;
    if (live) p.delta[(size_t)bh * N + i] = delta;
  }
  int yi = 0, xi = 0;
  if (BIAS == VRR_BIAS_POLY) {
    const int pi = ic > 0 ? ic - 1 : 0;
    yi = pi % p.bias_grid;
    xi = pi / p.bias_grid;
  }

  const uint64_t desc_q = smem_desc_sw128(smem_u32(sQ));
  const uint64_t desc_g = smem_desc_sw128(smem_u32(sG));
  constexpr uint32_t idesc_s = idesc_bf16(kRowsCta, kTile, 0, 0);  // [128 x 64] = A(K-major) B(K-major)
  constexpr uint32_t idesc_q = idesc_bf16(kRowsCta, kDh, 0, 1);    // dQ = dS(TMEM) . K (MN-major)

  float carry = 0.f;                         // TABLE: diagonal partial sum travelling through the warp
  float pw0 = 0.f, pw1 = 0.f, pw2 = 0.f, pw3 = 0.f;  // POLY: sum_j dS * d^k

  for (int t = 0; t < ntiles; ++t) {
    const int stage = t % kRing;
    const uint32_t use_parity = (uint32_t)((t / kRing) & 1);
    uint8_t* st = sRing + stage * 2 * kTile64Bytes;
    const int nvalid = min(kTile, N - t * kTile);

    if (tid == 0) {
      if (t == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_full[stage], use_parity);
      tc_fence_after();
      const uint64_t desc_k = smem_desc_sw128(smem_u32(st));
      const uint64_t desc_v = smem_desc_sw128(smem_u32(st + kTile64Bytes));
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ss(tmem_base + 0, desc_q + 2 * k, desc_k + 2 * k, idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ss(tmem_base + 64, desc_g + 2 * k, desc_v + 2 * k, idesc_s, k > 0);
      mma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, (uint32_t)(t & 1));
    tc_fence_after();

    if (warp_active) {
      uint32_t sr[64], dr[64];
      tmem_ld32(tmem_row + 0, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld32(tmem_row + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tmem_ld32(tmem_row + 64, *reinterpret_cast<uint32_t(*)[32]>(&dr[0]));
      tmem_ld32(tmem_row + 96, *reinterpret_cast<uint32_t(*)[32]>(&dr[32]));
      tmem_wait_ld();
      uint32_t packed[32];
#pragma unroll
      for (int e = 0; e < kTile; ++e) {
        const int j = t * kTile + e;
        float x = fmaf(__uint_as_float(sr[e]), p.scale_log2, -lse2);
        float dist_f = 0.f;
        if (BIAS == VRR_BIAS_TABLE) {
          x += lut[min(max(ic - j + N - 1, 0), 2 * N - 2)];
        } else if (BIAS == VRR_BIAS_POLY) {
          const int yx = key_yx[min(j, N - 1)];
          const int dist = abs(yi - (yx >> 8)) + abs(xi - (yx & 255));
          dist_f = (float)dist;
          x += (ic == 0 || j == 0) ? 0.f : lut[dist];
        }
        const float pr = (live && e < nvalid) ? ex2(x) : 0.f;
        const float ds = pr * (__uint_as_float(dr[e]) - delta);
        if (BIAS == VRR_BIAS_TABLE) {
          // element (i, j) lies on diagonal i - j; lane l-1 handled the same diagonal one column ago
          const float up = __shfl_up_sync(0xffffffffu, carry, 1);
          carry = ds + (lane > 0 ? up : 0.f);
          if (lane == 31) atomicAdd(&hist[min(max(i - j + N - 1, 0), 2 * N - 2)], carry);  // true row index: the diagonal id
        } else if (BIAS == VRR_BIAS_POLY) {
          const float w = (ic == 0 || j == 0) ? 0.f : ds;
          pw0 += w;
          const float w1 = w * dist_f;
          pw1 += w1;
          const float w2 = w1 * dist_f;
          pw2 += w2;
          pw3 = fmaf(w2, dist_f, pw3);
        }
        if (e & 1) packed[e >> 1] = pack_bf16(__uint_as_float(sr[e - 1]), ds);
        else sr[e] = __float_as_uint(ds);  // park the even element until its odd partner arrives
      }
      tmem_st32(tmem_row + 0, packed);  // dS (bf16 pairs) over the S columns
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      const uint64_t desc_kmn = smem_desc_sw128(smem_u32(st));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        mma_ts(tmem_base + 128, tmem_base + 0 + kk * 8, desc_kmn + 128 * kk, idesc_q, (t | kk) != 0);
      mma_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, (uint32_t)(t & 1));
    tc_fence_after();
    if (tid == 0 && t + kRing < ntiles) {
      const int tn = t + kRing;
      mbar_expect_tx(&bar_full[stage], 2 * kTile64Bytes);
      tma_load_2d(st, &tmap_pl, &bar_full[stage], 0, BHN + bh * N + tn * kTile);
      tma_load_2d(st + kTile64Bytes, &tmap_pl, &bar_full[stage], 0, 2 * BHN + bh * N + tn * kTile);
    }
    __syncwarp();
  }

  // ---- dQ epilogue ---------------------------------------------------------------------------------
  if (warp_active) {
    uint32_t lo[32], hi[32];
    tmem_ld32(tmem_row + 128, lo);
    tmem_ld32(tmem_row + 160, hi);
    tmem_wait_ld();
    if (live) store_row_bf16(p.d_planes + ((size_t)bh * N + i) * kDh, lo, hi, p.scale);
  }

  // ---- bias-gradient flush ---------------------------------------------------------------------------
  if (BIAS == VRR_BIAS_TABLE) {
    // carries still travelling: lane l holds the partial sum of diagonal (i - j_last)
    const int j_last = ntiles * kTile - 1;
    if (warp_active && lane != 31) atomicAdd(&hist[min(max(i - j_last + N - 1, 0), 2 * N - 2)], carry);
    __syncthreads();
    float* dst = p.d_bias_param + (size_t)h * (2 * N - 1);
    for (int t = tid; t < 2 * N - 1; t += 128) {
      const float v = hist[t];
      if (v != 0.f) atomicAdd(dst + t, v);
    }
  } else if (BIAS == VRR_BIAS_POLY) {
    float v[4] = {pw0, pw1, pw2, pw3};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if (lane == 0) {
      float* dst = p.d_bias_param + (size_t)(p.bias_heads == 1 ? 0 : h) * p.bias_len;
      for (int k = 0; k < p.bias_len && k < 4; ++k) atomicAdd(dst + k, v[k]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// ============================================================================================ dK, dV
template <int BIAS>
__global__ void __launch_bounds__(128, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmap_pl, const __grid_constant__ CUtensorMap tmap_do,
                       const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                        // this CTA's 128 key rows
  uint8_t* sV = sK + kTile128Bytes;
  uint8_t* sRing = sV + kTile128Bytes;       // [kRing][Q tile 8 KB | dO tile 8 KB]
  const int npad = ((p.N + kTile - 1) / kTile) * kTile;
  float2* stats = reinterpret_cast<float2*>(sRing + kRing * 2 * kTile64Bytes);  // [npad] (lse*log2e, delta)
  float* lut = reinterpret_cast<float*>(stats + npad);
  uint16_t* key_yx = reinterpret_cast<uint16_t*>(lut + p.lut_floats);
  const int key_yx_bytes = (BIAS == VRR_BIAS_POLY) ? ((p.N * 2 + 15) & ~15) : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(key_yx) + key_yx_bytes);
  uint64_t* bar_full = bars;
  uint64_t* bar_kv = bars + kRing;
  uint64_t* bar_s = bars + kRing + 1;
  uint64_t* bar_o = bars + kRing + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kRing + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, H = p.H;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int j0 = blockIdx.x * kRowsCta;
  const int j = j0 + tid;
  const int jc = min(j, N - 1);
  const bool live = j < N;
  const bool warp_active = (j0 + warp * 32) < N;
  const int ntiles = (N + kTile - 1) / kTile;
  const int BHN = p.B * H * N;

  if (tid == 0) {
    for (int s = 0; s < kRing; ++s) mbar_init(&bar_full[s], 1);
    mbar_init(bar_kv, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  fill_lut<BIAS>(lut, key_yx, p, h);
  for (int t = tid; t < npad; t += 128) {
    float2 v;
    v.x = t < N ? p.lse[(size_t)bh * N + t] * kLog2e : INFINITY;  // +inf -> P = 0 for padded queries
    v.y = t < N ? p.delta[(size_t)bh * N + t] : 0.f;
    stats[t] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);

  if (tid == 0) {
    tma_prefetch_desc(&tmap_pl);
    tma_prefetch_desc(&tmap_do);
    mbar_expect_tx(bar_kv, 2 * kTile128Bytes);
    tma_load_2d(sK, &tmap_pl, bar_kv, 0, BHN + bh * N + j0);
    tma_load_2d(sK + kTile64Bytes, &tmap_pl, bar_kv, 0, BHN + bh * N + j0 + kTile);
    tma_load_2d(sV, &tmap_pl, bar_kv, 0, 2 * BHN + bh * N + j0);
    tma_load_2d(sV + kTile64Bytes, &tmap_pl, bar_kv, 0, 2 * BHN + bh * N + j0 + kTile);
    for (int s = 0; s < kRing && s < ntiles; ++s) {
      uint8_t* st = sRing + s * 2 * kTile64Bytes;
      mbar_expect_tx(&bar_full[s], 2 * kTile64Bytes);
      tma_load_2d(st, &tmap_pl, &bar_full[s], 0, bh * N + s * kTile);
      tma_load_2d(st + kTile64Bytes, &tmap_do, &bar_full[s], h * kDh, b * N + s * kTile);
    }
  }
  __syncwarp();

  int yj = 0, xj = 0;
  if (BIAS == VRR_BIAS_POLY) {
    const int pj = jc > 0 ? jc - 1 : 0;
    yj = pj % p.bias_grid;
    xj = pj / p.bias_grid;
  }
  const uint64_t desc_k = smem_desc_sw128(smem_u32(sK));
  const uint64_t desc_v = smem_desc_sw128(smem_u32(sV));
  constexpr uint32_t idesc_s = idesc_bf16(kRowsCta, kTile, 0, 0);
  constexpr uint32_t idesc_acc = idesc_bf16(kRowsCta, kDh, 0, 1);

  for (int t = 0; t < ntiles; ++t) {
    const int stage = t % kRing;
    const uint32_t use_parity = (uint32_t)((t / kRing) & 1);
    uint8_t* st = sRing + stage * 2 * kTile64Bytes;

    if (tid == 0) {
      if (t == 0) mbar_wait(bar_kv, 0);
      mbar_wait(&bar_full[stage], use_parity);
      tc_fence_after();
      const uint64_t desc_q = smem_desc_sw128(smem_u32(st));
      const uint64_t desc_g = smem_desc_sw128(smem_u32(st + kTile64Bytes));
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ss(tmem_base + 0, desc_k + 2 * k, desc_q + 2 * k, idesc_s, k > 0);   // S^T
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ss(tmem_base + 64, desc_v + 2 * k, desc_g + 2 * k, idesc_s, k > 0);  // dP^T
      mma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, (uint32_t)(t & 1));
    tc_fence_after();

    if (warp_active) {
      uint32_t sr[64], dr[64];
      tmem_ld32(tmem_row + 0, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld32(tmem_row + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tmem_ld32(tmem_row + 64, *reinterpret_cast<uint32_t(*)[32]>(&dr[0]));
      tmem_ld32(tmem_row + 96, *reinterpret_cast<uint32_t(*)[32]>(&dr[32]));
      tmem_wait_ld();
      uint32_t pp[32], pd[32];
#pragma unroll
      for (int e = 0; e < kTile; ++e) {
        const int i = t * kTile + e;  // query index = accumulator column
        const float2 stt = stats[i];
        float x = fmaf(__uint_as_float(sr[e]), p.scale_log2, -stt.x);
        if (BIAS == VRR_BIAS_TABLE) {
          x += lut[min(max(i - jc + N - 1, 0), 2 * N - 2)];
        } else if (BIAS == VRR_BIAS_POLY) {
          const int yx = key_yx[min(i, N - 1)];
          const int dist = abs(yj - (yx >> 8)) + abs(xj - (yx & 255));
          x += (i == 0 || jc == 0) ? 0.f : lut[dist];
        }
        const float pr = ex2(x);
        const float ds = pr * (__uint_as_float(dr[e]) - stt.y);
        if (e & 1) {
          pp[e >> 1] = pack_bf16(__uint_as_float(sr[e - 1]), pr);
          pd[e >> 1] = pack_bf16(__uint_as_float(dr[e - 1]), ds);
        } else {
          sr[e] = __float_as_uint(pr);
          dr[e] = __float_as_uint(ds);
        }
      }
      tmem_st32(tmem_row + 0, pp);   // P^T  (bf16 pairs) over S^T
      tmem_st32(tmem_row + 64, pd);  // dS^T (bf16 pairs) over dP^T
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      const uint64_t desc_qmn = smem_desc_sw128(smem_u32(st));
      const uint64_t desc_gmn = smem_desc_sw128(smem_u32(st + kTile64Bytes));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        mma_ts(tmem_base + 128, tmem_base + 0 + kk * 8, desc_gmn + 128 * kk, idesc_acc, (t | kk) != 0);   // dV
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        mma_ts(tmem_base + 192, tmem_base + 64 + kk * 8, desc_qmn + 128 * kk, idesc_acc, (t | kk) != 0);  // dK
      mma_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, (uint32_t)(t & 1));
    tc_fence_after();
    if (tid == 0 && t + kRing < ntiles) {
      const int tn = t + kRing;
      mbar_expect_tx(&bar_full[stage], 2 * kTile64Bytes);
      tma_load_2d(st, &tmap_pl, &bar_full[stage], 0, bh * N + tn * kTile);
      tma_load_2d(st + kTile64Bytes, &tmap_do, &bar_full[stage], h * kDh, b * N + tn * kTile);
    }
    __syncwarp();
  }

  if (warp_active) {
    uint32_t lo[32], hi[32];
    const size_t plane = (size_t)BHN * kDh;
    tmem_ld32(tmem_row + 128, lo);
    tmem_ld32(tmem_row + 160, hi);
    tmem_wait_ld();
    if (live) store_row_bf16(p.d_planes + 2 * plane + ((size_t)bh * N + j) * kDh, lo, hi, 1.f);
    __syncwarp();
    tmem_ld32(tmem_row + 192, lo);
    tmem_ld32(tmem_row + 224, hi);
    tmem_wait_ld();
    if (live) store_row_bf16(p.d_planes + plane + ((size_t)bh * N + j) * kDh, lo, hi, p.scale);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

size_t bwd_smem_bytes(int N, const vrr_bias_desc* bias, int* lut_floats, bool dkv) {
  int lf = 0;
  const int mode = bias ? bias->mode : VRR_BIAS_NONE;
  if (mode == VRR_BIAS_TABLE) lf = 2 * N - 1;
  else if (mode == VRR_BIAS_POLY) lf = 2 * bias->grid - 1;
  lf = (lf + 3) & ~3;
  if (lut_floats) *lut_floats = lf;
  size_t extra = (size_t)lf * 4;
  if (mode == VRR_BIAS_POLY) extra += (size_t)((N * 2 + 15) & ~15);
  if (dkv) extra += (size_t)(((N + kTile - 1) / kTile) * kTile) * 8;
  else if (mode == VRR_BIAS_TABLE) extra += (size_t)((2 * N - 1 + 3) & ~3) * 4;
  return 1024 + 2 * kTile128Bytes + (size_t)kRing * 2 * kTile64Bytes + extra + (kRing + 3) * 8 + 16;
}

}  // namespace

bool attn_bwd_tc_supported(int B, int H, int N, int Dh, const vrr_bias_desc* bias) {
  if (Dh != kDh || N < 1) return false;
  if ((long long)3 * B * H * N >= (1ll << 31)) return false;
  const int mode = bias ? bias->mode : VRR_BIAS_NONE;
  if (mode == VRR_BIAS_POLY && (bias->grid > 255 || bias->len > 4)) return false;  // power sums up to degree 3
  return bwd_smem_bytes(N, bias, nullptr, false) <= 110 * 1024 && bwd_smem_bytes(N, bias, nullptr, true) <= 110 * 1024;
}

int attn_bwd_tc(const void* planes, const vrr_bias_desc* bias, const void* out, const void* d_out, const float* lse,
                void* d_planes, float* d_bias_param, float* delta, int B, int H, int N, int Dh, float scale,
                cudaStream_t st) {
  (void)Dh;
  VRR_REQUIRE(((uintptr_t)planes & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)d_out & 15) == 0 &&
                  ((uintptr_t)d_planes & 15) == 0,
              VRR_ERR_INVALID_ARG, "attn_bwd (tcgen05): planes / out / d_out / d_planes must be 16-byte aligned");
  const int E = H * kDh;
  CUtensorMap tm_pl, tm_do;
  if (int rc = make_tmap_bf16(&tm_pl, planes, (uint64_t)3 * B * H * N, kDh, kDh * 2, kTile)) return rc;
  if (int rc = make_tmap_bf16(&tm_do, d_out, (uint64_t)B * N, (uint64_t)E, (uint64_t)E * 2, kTile)) return rc;
  BwdParams p;
  p.out = (const __nv_bfloat16*)out;
  p.d_out = (const __nv_bfloat16*)d_out;
  p.lse = lse;
  p.delta = delta;
  p.d_planes = (__nv_bfloat16*)d_planes;
  p.B = B; p.H = H; p.N = N;
  p.scale = scale;
  p.scale_log2 = scale * kLog2e;
  const int mode = bias ? bias->mode : VRR_BIAS_NONE;
  p.bias_param = mode != VRR_BIAS_NONE ? bias->param : nullptr;
  p.d_bias_param = d_bias_param;
  p.bias_heads = mode != VRR_BIAS_NONE ? bias->heads : 0;
  p.bias_len = mode != VRR_BIAS_NONE ? bias->len : 0;
  p.bias_grid = mode != VRR_BIAS_NONE ? bias->grid : 0;
  if (mode != VRR_BIAS_NONE)
    VRR_CUDA(cudaMemsetAsync(d_bias_param, 0, (size_t)bias->heads * bias->len * sizeof(float), st));
  dim3 grid(ceil_div(N, kRowsCta), B * H);
  const size_t smem_q = bwd_smem_bytes(N, bias, &p.lut_floats, false);
  const size_t smem_kv = bwd_smem_bytes(N, bias, &p.lut_floats, true);
#define LAUNCH(MODE)                                                                                            \
  do {                                                                                                          \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      VRR_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                    110 * 1024));                                                               \
      VRR_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                    110 * 1024));                                                               \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    attn_bwd_dq_tc_kernel<MODE><<<grid, 128, smem_q, st>>>(tm_pl, tm_do, p);                                    \
    VRR_LAUNCHED();                                                                                             \
    attn_bwd_dkv_tc_kernel<MODE><<<grid, 128, smem_kv, st>>>(tm_pl, tm_do, p);                                  \
    VRR_LAUNCHED();                                                                                             \
  } while (0)
  if (mode == VRR_BIAS_TABLE) LAUNCH(VRR_BIAS_TABLE);
  else if (mode == VRR_BIAS_POLY) LAUNCH(VRR_BIAS_POLY);
  else LAUNCH(VRR_BIAS_NONE);
#undef LAUNCH
  return VRR_OK;
}

}  // namespace vrr
