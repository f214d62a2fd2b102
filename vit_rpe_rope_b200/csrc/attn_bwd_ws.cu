// Fused attention backward for SHORT sequences (N <= 256), one persistent kernel: the whole-sequence variant.
// Autograd of models/vit.py:71-88 of the reference:
//   P = softmax(S), dV = P^T dO, dP = dO V^T, dS = P o (dP - delta), dQ = scale dS K, dK = scale dS^T Q.
//
// Variant 2 (attn_bwd_tc2.cu) is two kernels (dQ; dK/dV) that each read Q, K, V, dO from HBM (1.44x the algorithmic
// traffic, measured) and walk the sequence in 32-row tiles: ten small tcgen05.mma, four mbarrier round trips and a
// CTA prologue per tile.  Here
//   * one CTA per SM loops over work items (image, head); Q, K, V, dO of the NEXT item are prefetched by TMA into the
//     other half of shared memory (2 x 4 x N x 128 B) while the current one is processed: every operand is read from
//     HBM exactly once per launch;
//   * an item is a sequence of PASSES over 128 TMEM lanes against ALL columns at once:
//       dQ pass (lane = query row i, one per 128 rows):  S = Q K^T, dP = dO V^T  ->  dS  ->  dQ = dS K
//       dKV pass (lane = key j, one per 128 keys):       S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T  ->  dV, dK
//     i.e. 8 + ~26 large tcgen05.mma and ONE softmax-side round trip per pass instead of one per 32 columns;
//   * the columns of a pass are split into two halves owned by the two compute warpgroups (the big half
//     alternates), each half with its own "S ready" / "P ready" barriers, so one group's exponentials overlap the
//     other's MMAs; P / dS are written back in place (bf16) and consumed from TMEM by the accumulating MMAs;
//   * TMEM (512 columns): S at [0,256), dP at [256,512); the accumulators (dQ, or dK and dV) live in columns
//     [64,128) / [320,384) of the first half - S / dP columns that half's owner has already consumed;
//   * delta = rowsum(dO o O) and lse*log2(e) of the NEXT item are computed by a helper warp into shared memory
//     while the current item runs (rows past N get lse = +inf -> P = 0).
// Bias modes (relative table / polynomial) stay on variant 2 for now.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrr {

using namespace tc;

namespace {

constexpr int kDh = 64;
constexpr int kThreads = 384;   // warps 0-3 / 4-7 compute groups, 8 producer, 9 issuer, 10 statistics helper, 11 idle
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColDP = 256, kAcc1 = 64, kAcc2 = 320;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kNumBars = 12;
constexpr int kSmemMax = 232448;

struct BwsParams {
  const __nv_bfloat16* out;
  const float* lse;
  __nv_bfloat16* d_planes;
  int B, H, N;
  float scale, scale_log2;
  int total_items;
  long long* dbg;
};

__device__ __forceinline__ void dbg_stamp(const BwsParams& p, int region, int idx) {
  if (p.dbg != nullptr && blockIdx.x == 0 && idx < 256) p.dbg[region * 256 + idx] = clock64();
}

__device__ __forceinline__ void store_half_row(__nv_bfloat16* dst, const uint32_t (&v)[32], float mul) {
#pragma unroll
  for (int v8 = 0; v8 < 4; ++v8) {
    uint4 w;
    w.x = pack_bf16(__uint_as_float(v[v8 * 8 + 0]) * mul, __uint_as_float(v[v8 * 8 + 1]) * mul);
    w.y = pack_bf16(__uint_as_float(v[v8 * 8 + 2]) * mul, __uint_as_float(v[v8 * 8 + 3]) * mul);
    w.z = pack_bf16(__uint_as_float(v[v8 * 8 + 4]) * mul, __uint_as_float(v[v8 * 8 + 5]) * mul);
    w.w = pack_bf16(__uint_as_float(v[v8 * 8 + 6]) * mul, __uint_as_float(v[v8 * 8 + 7]) * mul);
    *reinterpret_cast<uint4*>(dst + v8 * 8) = w;
  }
}

// dQ pass, W (32 or 16) columns = keys j0 .. j0+W-1 of this thread's query row:
//   dS = exp2(S c - lse2) (dP - delta), keys past N masked; dS (bf16 pairs) over the S columns.
template <int W>
__device__ __forceinline__ void dq_chunk(const BwsParams& p, uint32_t t_s, uint32_t t_dp, uint32_t t_dst, int j0,
                                         float neg_lse2, float delta) {
  uint32_t s[W], d[W], packed[W / 2];
  if constexpr (W == 32) {
    tmem_ld32(t_s, s);
    tmem_ld32(t_dp, d);
  } else {
    tmem_ld16(t_s, s);
    tmem_ld16(t_dp, d);
  }
  tmem_wait_ld();
  const bool full = j0 + W <= p.N;
#pragma unroll
  for (int e = 0; e < W; e += 2) {
    float p0 = ex2(fmaf(__uint_as_float(s[e]), p.scale_log2, neg_lse2));
    float p1 = ex2(fmaf(__uint_as_float(s[e + 1]), p.scale_log2, neg_lse2));
    if (!full) {
      p0 = j0 + e < p.N ? p0 : 0.f;
      p1 = j0 + e + 1 < p.N ? p1 : 0.f;
    }
    const float ds0 = p0 * (__uint_as_float(d[e]) - delta);
    const float ds1 = p1 * (__uint_as_float(d[e + 1]) - delta);
    packed[e >> 1] = pack_bf16(ds0, ds1);
  }
  if constexpr (W == 32) tmem_st16(t_dst, packed);
  else tmem_st8(t_dst, packed);
}

// dK/dV pass, W columns = query rows i0 .. i0+W-1 of this thread's key: per-column (lse2, delta) from shared memory;
//   P^T (bf16) over the S^T columns, dS^T (bf16) over the dP^T columns.
template <int W>
__device__ __forceinline__ void dkv_chunk(const BwsParams& p, uint32_t t_s, uint32_t t_dp, uint32_t t_dst_p,
                                          uint32_t t_dst_ds, const float4* stats2) {
  uint32_t s[W], d[W], pp[W / 2], pd[W / 2];
  if constexpr (W == 32) {
    tmem_ld32(t_s, s);
    tmem_ld32(t_dp, d);
  } else {
    tmem_ld16(t_s, s);
    tmem_ld16(t_dp, d);
  }
  tmem_wait_ld();
#pragma unroll
  for (int e = 0; e < W; e += 2) {
    const float4 st = stats2[e >> 1];  // (lse2, delta) of rows i0+e and i0+e+1; lse2 = +inf past N -> P = 0
    const float p0 = ex2(fmaf(__uint_as_float(s[e]), p.scale_log2, -st.x));
    const float p1 = ex2(fmaf(__uint_as_float(s[e + 1]), p.scale_log2, -st.z));
    const float ds0 = p0 * (__uint_as_float(d[e]) - st.y);
    const float ds1 = p1 * (__uint_as_float(d[e + 1]) - st.w);
    pp[e >> 1] = pack_bf16(p0, p1);
    pd[e >> 1] = pack_bf16(ds0, ds1);
  }
  if constexpr (W == 32) {
    tmem_st16(t_dst_p, pp);
    tmem_st16(t_dst_ds, pd);
  } else {
    tmem_st8(t_dst_p, pp);
    tmem_st8(t_dst_ds, pd);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_pl64, const __grid_constant__ CUtensorMap tm_pl16,
                   const __grid_constant__ CUtensorMap tm_do64, const __grid_constant__ CUtensorMap tm_do16,
                   const __grid_constant__ BwsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = p.N, H = p.H, E = H * kDh;
  const int npad = (N + 15) & ~15;
  const int tbytes = npad * 128;                   // one tensor of one item (multiple of 2048)
  const int slot_bytes = 4 * tbytes;               // Q | K | V | dO
  float2* stats = reinterpret_cast<float2*>(smem + 2 * slot_bytes);   // [2 items][256] (lse * log2e, delta)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 512);
  uint64_t* bar_full = bars;              // [2] item data landed
  uint64_t* bar_empty = bars + 2;         // [2] last MMA of the item retired -> slot reusable
  uint64_t* bar_sfull = bars + 4;         // [2 halves] S and dP of the half ready
  uint64_t* bar_pfull = bars + 6;         // [2 halves] P / dS of the half stored (128 arrivals)
  uint64_t* bar_accfull = bars + 8;       // accumulators of the pass ready
  uint64_t* bar_accempty = bars + 9;      // epilogue of the pass has read them (256 arrivals)
  uint64_t* bar_stfull = bars + 10;       // [2] statistics of the item written (32 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = p.total_items;
  const int BHN = p.B * H * N;
  const int ntile = (N + 127) >> 7;                // 128-lane tiles of rows / of keys
  const int cA = npad < 128 ? npad : 128;          // columns of the first half of a pass
  const int cB = npad - cA;                        // columns of the second half (0: one group idles)
  const int npass = 2 * ntile;                     // passes per item: ntile dQ passes, then ntile dK/dV passes

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
      mbar_init(&bar_sfull[s], 1);
      mbar_init(&bar_pfull[s], 128);
      mbar_init(&bar_stfull[s], 32);
    }
    mbar_init(bar_accfull, 1);
    mbar_init(bar_accempty, 256);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 8) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ============================================ TMA producer ============================================
    if (elect_one()) {
      tma_prefetch_desc(&tm_pl64);
      tma_prefetch_desc(&tm_pl16);
      tma_prefetch_desc(&tm_do64);
      tma_prefetch_desc(&tm_do16);
    }
    __syncwarp();
    const int nb64 = npad >> 6, nb16 = (npad & 63) >> 4;
    int k = 0;
    for (int bh = blockIdx.x; bh < total; bh += gridDim.x, ++k) {
      const int sl = k & 1;
      const int b = bh / H, h = bh - b * H;
      mbar_wait(&bar_empty[sl], (uint32_t)(((k >> 1) & 1) ^ 1));
      if (elect_one()) {
        uint8_t* base = smem + sl * slot_bytes;
        mbar_expect_tx(&bar_full[sl], (uint32_t)slot_bytes);
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {  // Q, K, V planes
          uint8_t* dst = base + t * tbytes;
          const int row0 = t * BHN + bh * N;
          for (int j = 0; j < nb64; ++j) tma_load_2d(dst + j * 8192, &tm_pl64, &bar_full[sl], 0, row0 + j * 64);
          for (int j = 0; j < nb16; ++j)
            tma_load_2d(dst + nb64 * 8192 + j * 2048, &tm_pl16, &bar_full[sl], 0, row0 + nb64 * 64 + j * 16);
        }
        uint8_t* dst = base + 3 * tbytes;  // dO rows of image b, columns of head h
        for (int j = 0; j < nb64; ++j) tma_load_2d(dst + j * 8192, &tm_do64, &bar_full[sl], h * kDh, b * N + j * 64);
        for (int j = 0; j < nb16; ++j)
          tma_load_2d(dst + nb64 * 8192 + j * 2048, &tm_do16, &bar_full[sl], h * kDh, b * N + nb64 * 64 + j * 16);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ============================================ MMA issuer ==============================================
    const uint32_t smem_b = smem_u32(smem);
    uint32_t n = 0;  // global pass counter of this CTA
    int k = 0;
    for (int bh = blockIdx.x; bh < total; bh += gridDim.x, ++k) {
      const int sl = k & 1;
      const uint32_t q_b = smem_b + sl * slot_bytes, k_b = q_b + tbytes, v_b = k_b + tbytes, g_b = v_b + tbytes;
      mbar_wait(&bar_full[sl], (uint32_t)((k >> 1) & 1));
      for (int ps = 0; ps < npass; ++ps, ++n) {
        const bool is_dq = ps < ntile;
        const int tile = is_dq ? ps : ps - ntile;
        const uint32_t par = n & 1;
        // lanes: dQ pass -> Q / dO rows of the tile against K / V;  dK/dV pass -> K / V rows against Q / dO
        const uint32_t a_s = (is_dq ? q_b : k_b) + tile * 16384, a_p = (is_dq ? g_b : v_b) + tile * 16384;
        const uint32_t b_s = is_dq ? k_b : q_b, b_p = is_dq ? v_b : g_b;
        mbar_wait(bar_accempty, par ^ 1);  // the previous pass' epilogue has read its accumulators
        tc_fence_after();
        if (lane == 0) dbg_stamp(p, 2, (int)n * 8 + 0);
        if (elect_one()) {
          const uint64_t das = smem_desc_sw128(a_s), dap = smem_desc_sw128(a_p);
          {
            const uint32_t idesc = idesc_bf16(128, cA, 0, 0);
            const uint64_t dbs = smem_desc_sw128(b_s), dbp = smem_desc_sw128(b_p);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_ss(tmem_base, das + 2 * kk, dbs + 2 * kk, idesc, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_ss(tmem_base + kColDP, dap + 2 * kk, dbp + 2 * kk, idesc, kk > 0);
            mma_commit(&bar_sfull[0]);
          }
          if (cB > 0) {
            const uint32_t idesc = idesc_bf16(128, cB, 0, 0);
            const uint64_t dbs = smem_desc_sw128(b_s + cA * 128), dbp = smem_desc_sw128(b_p + cA * 128);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_ss(tmem_base + cA, das + 2 * kk, dbs + 2 * kk, idesc, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_ss(tmem_base + kColDP + cA, dap + 2 * kk, dbp + 2 * kk, idesc, kk > 0);
            mma_commit(&bar_sfull[1]);
          }
        }
        __syncwarp();
        if (lane == 0) dbg_stamp(p, 2, (int)n * 8 + 1);
        constexpr uint32_t idesc_acc = idesc_bf16(128, kDh, 0, 1);
        // accumulating MMAs of one half: A = P / dS (bf16) from TMEM, B = the column-side operand, MN-major
        for (int half = 0; half < 2; ++half) {
          const int c0 = half == 0 ? 0 : cA, cw = half == 0 ? cA : cB;
          if (cw == 0) break;
          mbar_wait(&bar_pfull[half], par);
          tc_fence_after();
          if (lane == 0) dbg_stamp(p, 2, (int)n * 8 + 2 + 2 * half);
          if (elect_one()) {
            const int ksteps = cw >> 4;
            if (is_dq) {
              const uint64_t dk = smem_desc_sw128(k_b) + (uint64_t)(128 * (c0 >> 4));
              for (int kk = 0; kk < ksteps; ++kk)
                mma_ts(tmem_base + kAcc1, tmem_base + c0 + kk * 8, dk + 128 * kk, idesc_acc, (half | kk) != 0);
            } else {
              const uint64_t dg = smem_desc_sw128(g_b) + (uint64_t)(128 * (c0 >> 4));
              const uint64_t dq = smem_desc_sw128(q_b) + (uint64_t)(128 * (c0 >> 4));
              for (int kk = 0; kk < ksteps; ++kk)  // dV += P^T dO
                mma_ts(tmem_base + kAcc2, tmem_base + c0 + kk * 8, dg + 128 * kk, idesc_acc, (half | kk) != 0);
              for (int kk = 0; kk < ksteps; ++kk)  // dK += dS^T Q
                mma_ts(tmem_base + kAcc1, tmem_base + kColDP + c0 + kk * 8, dq + 128 * kk, idesc_acc, (half | kk) != 0);
            }
            if (half == 1 || cB == 0) {
              mma_commit(bar_accfull);
              if (ps == npass - 1) mma_commit(&bar_empty[sl]);
            }
          }
          __syncwarp();
          if (lane == 0) dbg_stamp(p, 2, (int)n * 8 + 3 + 2 * half);
        }
      }
    }
  } else if (warp == 10) {
    // ============================================ statistics helper =======================================
    // stats[k & 1][i] = (lse_i * log2 e, delta_i = sum_d dO[i][d] O[i][d]) for the rows of item k, one item ahead
    int k = 0;
    for (int bh = blockIdx.x; bh < total; bh += gridDim.x, ++k) {
      const int sl = k & 1;
      const int b = bh / H, h = bh - b * H;
      float2* st = stats + sl * 256;
      const uint8_t* sG = smem + sl * slot_bytes + 3 * tbytes;
      mbar_wait(&bar_full[sl], (uint32_t)((k >> 1) & 1));
#pragma unroll 1
      for (int r0 = 0; r0 < 256; r0 += 32) {
        const int i = r0 + lane;
        float2 v = make_float2(INFINITY, 0.f);
        if (i < N) {
          const uint4* o4 = reinterpret_cast<const uint4*>(p.out + ((size_t)b * N + i) * E + h * kDh);
          const uint8_t* grow = sG + i * 128;
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 ov = __ldg(o4 + c);
            const uint4 gv = *reinterpret_cast<const uint4*>(grow + ((c ^ (i & 7)) << 4));
            const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[e]));
              const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]));
              acc = fmaf(of.x, gf.x, acc);
              acc = fmaf(of.y, gf.y, acc);
            }
          }
          v = make_float2(p.lse[(size_t)bh * N + i] * kLog2e, acc);
        }
        st[i] = v;
      }
      mbar_arrive(&bar_stfull[sl]);
    }
  } else if (warp < 8) {
    // ============================================ compute groups ==========================================
    const int w = warp >> 2, wq = warp & 3;
    const int lrow = wq * 32 + lane;  // TMEM lane of this thread
    const uint32_t tmem_row = tmem_base + ((uint32_t)(wq * 32) << 16);
    const size_t plane = (size_t)BHN * kDh;
    uint32_t n = 0;
    int k = 0;
    for (int bh = blockIdx.x; bh < total; bh += gridDim.x, ++k) {
      const int sl = k & 1;
      const float2* st = stats + sl * 256;
      mbar_wait(&bar_stfull[sl], (uint32_t)((k >> 1) & 1));
      for (int ps = 0; ps < npass; ++ps, ++n) {
        const bool is_dq = ps < ntile;
        const int tile = is_dq ? ps : ps - ntile;
        const uint32_t par = n & 1;
        const int half = (w == (int)(n & 1)) ? 0 : 1;  // the big half alternates between the groups
        const int c0 = half == 0 ? 0 : cA, cw = half == 0 ? cA : cB;
        const int idx = tile * 128 + lrow;             // query row (dQ pass) / key (dK/dV pass) of this thread
        const bool warp_live = tile * 128 + wq * 32 < N;
        const bool dbg_on = p.dbg != nullptr && (tid & 127) == 0;
        if (cw > 0) {
          if (dbg_on) dbg_stamp(p, w, (int)n * 4 + 0);
          mbar_wait(&bar_sfull[half], par);
          tc_fence_after();
          if (dbg_on) dbg_stamp(p, w, (int)n * 4 + 1);
          if (warp_live) {
            const uint32_t t_s = tmem_row + c0, t_dp = tmem_row + kColDP + c0;
            const int n32 = cw >> 5;
            if (is_dq) {
              const float2 mine = st[min(idx, 255)];
              const float neg_lse2 = -mine.x, delta = mine.y;
#pragma unroll 1
              for (int c = 0; c < n32; ++c)
                dq_chunk<32>(p, t_s + c * 32, t_dp + c * 32, t_s + c * 16, c0 + c * 32, neg_lse2, delta);
              if (cw & 16) dq_chunk<16>(p, t_s + n32 * 32, t_dp + n32 * 32, t_s + n32 * 16, c0 + n32 * 32, neg_lse2, delta);
            } else {
              const float4* st2 = reinterpret_cast<const float4*>(st + c0);
#pragma unroll 1
              for (int c = 0; c < n32; ++c)
                dkv_chunk<32>(p, t_s + c * 32, t_dp + c * 32, t_s + c * 16, t_dp + c * 16, st2 + c * 16);
              if (cw & 16)
                dkv_chunk<16>(p, t_s + n32 * 32, t_dp + n32 * 32, t_s + n32 * 16, t_dp + n32 * 16, st2 + n32 * 16);
            }
            tmem_wait_st();
          }
          tc_fence_before();
          mbar_arrive(&bar_pfull[half]);
          if (dbg_on) dbg_stamp(p, w, (int)n * 4 + 2);
        }
        // ---- epilogue of the pass ----------------------------------------------------------------------
        mbar_wait(bar_accfull, par);
        tc_fence_after();
        if (dbg_on) dbg_stamp(p, w, (int)n * 4 + 3);
        uint32_t lo[32], hi[32];
        if (warp_live) {
          if (is_dq) {
            tmem_ld32(tmem_row + kAcc1 + w * 32, lo);  // group w stores channels [32 w, 32 w + 32) of dQ
          } else {
            tmem_ld32(tmem_row + (w == 0 ? kAcc2 : kAcc1), lo);  // group 0: dV, group 1: dK
            tmem_ld32(tmem_row + (w == 0 ? kAcc2 : kAcc1) + 32, hi);
          }
          tmem_wait_ld();
        }
        tc_fence_before();
        mbar_arrive(bar_accempty);
        if (warp_live && idx < N) {
          if (is_dq) {
            store_half_row(p.d_planes + ((size_t)bh * N + idx) * kDh + w * 32, lo, p.scale);
          } else {
            __nv_bfloat16* dst = p.d_planes + (w == 0 ? 2 : 1) * plane + ((size_t)bh * N + idx) * kDh;
            const float mul = w == 0 ? 1.f : p.scale;
            store_half_row(dst, lo, mul);
            store_half_row(dst + 32, hi, mul);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, kTmemCols);
}

// The 128-row A operand of a pass starts at row tile * 128 of a tensor that holds round16(N) rows: the tensor core
// reads up to 127 rows (16 KB) past it - harmless garbage for lanes that are never stored, but the bytes must exist:
// a 16 KB tail (holding the statistics and the barriers) follows the last slot.
size_t bws_smem_bytes(int N) {
  const int npad = (N + 15) & ~15;
  static_assert(512 * 8 + kNumBars * 8 + 16 <= 16384, "tail too small");
  return 1024 + (size_t)2 * 4 * npad * 128 + 16384;
}

std::atomic<long long*> g_bws_dbg{nullptr};

}  // namespace

void attn_bwd_ws_set_debug(long long* buf) { g_bws_dbg.store(buf); }

bool attn_bwd_ws_supported(int B, int H, int N, int Dh, const vrr_bias_desc* bias) {
  if (Dh != kDh || N < 1 || N > 256) return false;
  if (bias && bias->mode != VRR_BIAS_NONE) return false;
  if ((long long)3 * B * H * N >= (1ll << 31)) return false;
  return bws_smem_bytes(N) <= (size_t)kSmemMax;
}

int attn_bwd_ws(const void* planes, const void* out, const void* d_out, const float* lse, void* d_planes, int B, int H,
                int N, int Dh, float scale, cudaStream_t st) {
  (void)Dh;
  VRR_REQUIRE(((uintptr_t)planes & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)d_out & 15) == 0 &&
                  ((uintptr_t)d_planes & 15) == 0,
              VRR_ERR_INVALID_ARG, "attn_bwd (tcgen05): planes / out / d_out / d_planes must be 16-byte aligned");
  const int E = H * kDh;
  CUtensorMap tm_pl64, tm_pl16, tm_do64, tm_do16;
  if (int rc = make_tmap_2d(&tm_pl64, planes, 2, (uint64_t)3 * B * H * N, kDh, kDh * 2, 64, 64)) return rc;
  if (int rc = make_tmap_2d(&tm_pl16, planes, 2, (uint64_t)3 * B * H * N, kDh, kDh * 2, 16, 64)) return rc;
  if (int rc = make_tmap_2d(&tm_do64, d_out, 2, (uint64_t)B * N, (uint64_t)E, (uint64_t)E * 2, 64, 64)) return rc;
  if (int rc = make_tmap_2d(&tm_do16, d_out, 2, (uint64_t)B * N, (uint64_t)E, (uint64_t)E * 2, 16, 64)) return rc;
  BwsParams p;
  p.out = (const __nv_bfloat16*)out;
  p.lse = lse;
  p.d_planes = (__nv_bfloat16*)d_planes;
  p.B = B; p.H = H; p.N = N;
  p.scale = scale;
  p.scale_log2 = scale * kLog2e;
  p.total_items = B * H;
  p.dbg = g_bws_dbg.load();
  const size_t smem = bws_smem_bytes(N);
  const int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  VRR_SMEM_ATTR_ONCE(attn_bwd_ws_kernel, kSmemMax);
  attn_bwd_ws_kernel<<<grid, kThreads, smem, st>>>(tm_pl64, tm_pl16, tm_do64, tm_do16, p);
  VRR_LAUNCHED();
  return VRR_OK;
}

}  // namespace vrr
