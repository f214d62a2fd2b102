// Fused attention backward for SHORT sequences (N <= 200: the 65..197-token ViT configurations), one persistent
// kernel, second schedule: double-buffered column groups processed by all eight compute warps.
// Autograd of models/vit.py:71-88 of the reference:
//   P = softmax(S), dV = P^T dO, dP = dO V^T, dS = P o (dP - delta), dQ = scale dS K, dK = scale dS^T Q.
//
// Same decomposition as attn_bwd_ws.cu - one CTA per SM, work item = (image, head), Q / K / V / dO of the NEXT item
// prefetched by TMA (every operand read from HBM once), an item = dQ LANE TILES (lane = query row: S = Q K^T,
// dP = dO V^T -> dS -> dQ = dS K) followed by dKV lane tiles (lane = key: S^T, dP^T -> P^T, dS^T -> dV, dK) - but the
// clock64 timelines of that kernel showed each of its two streams alternating between "wait for the tensor core"
// and "compute" with little overlap, and per-thread row stores costing ~2 000 cycles per lane tile.  Here
//   * a lane tile is walked in COLUMN GROUPS of <= 96 columns through TWO TMEM buffers (S | dP, 192 columns each):
//     the issuer runs two groups ahead, so the S / dP MMAs of group g+1 and the accumulating MMAs of group g-1
//     execute while the threads are in the exponentials of group g - across lane tiles and items too;
//   * all eight compute warps work on the same group: warp w owns TMEM lanes 32 (w % 4) .. +31 and the column half
//     w / 4, writes P / dS (bf16) back in place over the start of ITS OWN columns, and one barrier per buffer
//     (256 arrivals) hands the group to the accumulating MMAs (A from TMEM in two runs, one per half);
//   * accumulators: dQ or dV at columns [384,448), dK at [448,512);
//   * epilogue: fp32 -> bf16 -> a 2 KB staging buffer per warp -> TMA store through a 3-D map of
//     d_planes[3 B H][N][64] (rows past N are clipped), 32 rows x 32 channels at a time;
//   * tensors are packed at round8(N) rows so that two items plus the staging buffers fit in 227 KB.
// delta = rowsum(dO o O) and lse * log2(e) of the next item come from a helper warp, as before.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrr {

using namespace tc;

namespace {

constexpr int kDh = 64;
constexpr int kThreads = 384;  // warps 0-7 compute, 8 producer, 9 issuer, 10 statistics helper, 11 idle
constexpr uint32_t kTmemCols = 512;
constexpr int kGW = 96;                       // columns per group
constexpr uint32_t kBufCols = 2 * kGW;        // S | dP
constexpr uint32_t kAccA = 384, kAccB = 448;  // dQ or dV | dK
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kSmemMax = 232448;
constexpr int kTailStats = 1024, kTailBars = 5120, kTailStage = 5376, kTailBytes = 22528;  // zero gap | stats | barriers | 8 x 2 KB staging

struct Bws2Params {
  const __nv_bfloat16* out;
  const float* lse;
  int B, H, N;
  float scale, scale_log2;
  int total_items;
  int tb;  // bytes of one tensor of one item in shared memory
  long long* dbg;
};

__device__ __forceinline__ void dbg_stamp(const Bws2Params& p, int region, int idx) {
  if (p.dbg != nullptr && blockIdx.x == 0 && idx < 256) p.dbg[region * 256 + idx] = clock64();
}

// 16 accumulator columns of one TMEM lane.
//   DQ  (lane = query row i, column = key j):  dS = exp2(S c - lse2_i) (dP - delta_i), keys past N masked;
//        dS (bf16 pairs) -> t_dst_ds.
//   !DQ (lane = key j, column = query row i):  per-column (lse2_i, delta_i) from shared memory (lse2 = +inf past N
//        -> P = 0);  P^T (bf16) -> t_dst_p, dS^T (bf16) -> t_dst_ds.
template <bool DQ>
__device__ __forceinline__ void process16(const Bws2Params& p, uint32_t t_s, uint32_t t_dp, uint32_t t_dst_p, uint32_t t_dst_ds,
                                          int col0, float neg_lse2, float delta, const float4* stats2) {
  uint32_t s[16], d[16], pp[8], pd[8];
  tmem_ld16(t_s, s);
  tmem_ld16(t_dp, d);
  tmem_wait_ld();
  if (DQ) {
    const bool full = col0 + 16 <= p.N;
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      float p0 = ex2(fmaf(__uint_as_float(s[e]), p.scale_log2, neg_lse2));
      float p1 = ex2(fmaf(__uint_as_float(s[e + 1]), p.scale_log2, neg_lse2));
      if (!full) {
        p0 = col0 + e < p.N ? p0 : 0.f;
        p1 = col0 + e + 1 < p.N ? p1 : 0.f;
      }
      pd[e >> 1] = pack_bf16(p0 * (__uint_as_float(d[e]) - delta), p1 * (__uint_as_float(d[e + 1]) - delta));
    }
    tmem_st8(t_dst_ds, pd);
  } else {
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
      const float4 st = stats2[e >> 1];  // (lse2, delta) of rows col0+e and col0+e+1
      const float p0 = ex2(fmaf(__uint_as_float(s[e]), p.scale_log2, -st.x));
      const float p1 = ex2(fmaf(__uint_as_float(s[e + 1]), p.scale_log2, -st.z));
      pp[e >> 1] = pack_bf16(p0, p1);
      pd[e >> 1] = pack_bf16(p0 * (__uint_as_float(d[e]) - st.y), p1 * (__uint_as_float(d[e + 1]) - st.w));
    }
    tmem_st8(t_dst_p, pp);
    tmem_st8(t_dst_ds, pd);
  }
}

// Walks the (item, lane tile, column group) sequence of one CTA; the issuer and the compute warps step through the
// same sequence.  Lane tiles of an item: ntile dQ tiles, then ntile dKV tiles.
struct Cursor {
  int k, bh, lt, gi;  // item ordinal of this CTA, item, lane tile, column group
  int stride, total, ntile, ngr;
  __device__ __forceinline__ bool valid() const { return bh < total; }
  __device__ __forceinline__ void advance() {
    if (++gi == ngr) {
      gi = 0;
      if (++lt == 2 * ntile) {
        lt = 0;
        ++k;
        bh += stride;
      }
    }
  }
  __device__ __forceinline__ bool is_dq() const { return lt < ntile; }
  __device__ __forceinline__ int tile() const { return lt < ntile ? lt : lt - ntile; }
  __device__ __forceinline__ bool first_of_item() const { return lt == 0 && gi == 0; }
  __device__ __forceinline__ bool last_of_tile() const { return gi == ngr - 1; }
  __device__ __forceinline__ bool last_of_item() const { return gi == ngr - 1 && lt == 2 * ntile - 1; }
};

// stage 32 rows x 32 channels (this thread = one row, 32 fp32 accumulators) as bf16 and store them with TMA
__device__ __forceinline__ void stage_and_store(const CUtensorMap* map, uint8_t* stage, bool leader, const uint32_t (&v)[32], float mul,
                                                int lane, int ch0, int row0, int plane) {
  if (leader) bulk_wait_read<0>();  // the previous store has finished reading the staging buffer
  __syncwarp();
  const uint32_t base = smem_u32(stage) + (uint32_t)lane * 64u;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    st_shared_v4(base + c * 16, pack_bf16(__uint_as_float(v[c * 8 + 0]) * mul, __uint_as_float(v[c * 8 + 1]) * mul),
                 pack_bf16(__uint_as_float(v[c * 8 + 2]) * mul, __uint_as_float(v[c * 8 + 3]) * mul),
                 pack_bf16(__uint_as_float(v[c * 8 + 4]) * mul, __uint_as_float(v[c * 8 + 5]) * mul),
                 pack_bf16(__uint_as_float(v[c * 8 + 6]) * mul, __uint_as_float(v[c * 8 + 7]) * mul));
  fence_proxy_async_smem();
  __syncwarp();
  if (leader) {
    tma_store_3d(map, stage, ch0, row0, plane);
    bulk_commit();
  }
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_ws2_kernel(const __grid_constant__ CUtensorMap tm_pl64, const __grid_constant__ CUtensorMap tm_plt,
                    const __grid_constant__ CUtensorMap tm_do64, const __grid_constant__ CUtensorMap tm_dot,
                    const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ Bws2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = p.N, H = p.H, E = H * kDh;
  const int npad = (N + 15) & ~15;                 // columns of a lane tile (MMA N granularity)
  const int rows8 = (N + 7) & ~7;                  // rows of a tensor that TMA delivers
  const int tb = p.tb, slot_bytes = 4 * tb;        // Q | K | V | dO
  uint8_t* tail = smem + 2 * slot_bytes;
  float2* stats = reinterpret_cast<float2*>(tail + kTailStats);   // [2 items][256] (lse * log2e, delta)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + kTailBars);
  uint64_t* bar_full = bars;            // [2] item data landed
  uint64_t* bar_empty = bars + 2;       // [2] last MMA of the item retired -> slot reusable
  uint64_t* bar_sfull = bars + 4;       // [2 buffers] S and dP of the group ready
  uint64_t* bar_pfull = bars + 6;       // [2 buffers] P / dS of the group stored (256 arrivals)
  uint64_t* bar_accfull = bars + 8;     // accumulators of the lane tile ready
  uint64_t* bar_accempty = bars + 9;    // epilogue has read them (256 arrivals)
  uint64_t* bar_stfull = bars + 10;     // [2] statistics of the item written (32 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = p.total_items;
  const int BHN = p.B * H * N;
  const int ntile = (N + 127) >> 7;
  const int ngr = (npad + kGW - 1) / kGW;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
      mbar_init(&bar_sfull[s], 1);
      mbar_init(&bar_pfull[s], 256);
      mbar_init(&bar_stfull[s], 32);
    }
    mbar_init(bar_accfull, 1);
    mbar_init(bar_accempty, 256);
    fence_mbar_init();
  }
  // Rows [round8(N), round16(N)) of a tensor are read by the MMAs (as columns whose P is exactly 0) but never loaded:
  // they alias the first rows of the NEXT tensor - finite data - except after the last tensor of a slot.  Those two
  // places (the start of slot 1 before its first load, and the gap that opens the tail) are zeroed once: 0 x NaN from
  // an uninitialised byte pattern (or from the +inf statistics that used to follow) would poison dK and dV.
  for (int i = tid; i < 256; i += kThreads) {
    reinterpret_cast<uint32_t*>(tail)[i] = 0u;
    reinterpret_cast<uint32_t*>(smem + slot_bytes)[i] = 0u;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 8) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // NB: with ~227 KB of shared memory the L1 that would absorb register spills is nearly gone; keep `ptxas -v` at
  // 0 spill bytes for this kernel.

  if (warp == 8) {
    // ============================================ TMA producer ============================================
    if (elect_one()) {
      tma_prefetch_desc(&tm_pl64);
      tma_prefetch_desc(&tm_plt);
      tma_prefetch_desc(&tm_do64);
      tma_prefetch_desc(&tm_dot);
      tma_prefetch_desc(&tm_out);
    }
    __syncwarp();
    const int nb64 = rows8 >> 6, trows = rows8 & 63;
    int k = 0;
    for (int bh = blockIdx.x; bh < total; bh += gridDim.x, ++k) {
      const int sl = k & 1;
      const int b = bh / H, h = bh - b * H;
      mbar_wait(&bar_empty[sl], (uint32_t)(((k >> 1) & 1) ^ 1));
      if (elect_one()) {
        uint8_t* base = smem + sl * slot_bytes;
        mbar_expect_tx(&bar_full[sl], (uint32_t)(4 * rows8 * 128));
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {  // Q, K, V planes
          uint8_t* dst = base + t * tb;
          const int row0 = t * BHN + bh * N;
          for (int j = 0; j < nb64; ++j) tma_load_2d(dst + j * 8192, &tm_pl64, &bar_full[sl], 0, row0 + j * 64);
          if (trows) tma_load_2d(dst + nb64 * 8192, &tm_plt, &bar_full[sl], 0, row0 + nb64 * 64);
        }
        uint8_t* dst = base + 3 * tb;  // dO rows of image b, columns of head h
        for (int j = 0; j < nb64; ++j) tma_load_2d(dst + j * 8192, &tm_do64, &bar_full[sl], h * kDh, b * N + j * 64);
        if (trows) tma_load_2d(dst + nb64 * 8192, &tm_dot, &bar_full[sl], h * kDh, b * N + nb64 * 64);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ============================================ MMA issuer ==============================================
    const uint32_t smem_b = smem_u32(smem);
    constexpr uint32_t idesc_acc = idesc_bf16(128, kDh, 0, 1);
    Cursor cs{0, (int)blockIdx.x, 0, 0, (int)gridDim.x, total, ntile, ngr};  // next group whose S / dP get issued
    Cursor ca = cs;                                                         // next group whose accumulating MMAs get issued
    uint32_t gs = 0, ga = 0, nt = 0;
    auto issue_s = [&](const Cursor& c, uint32_t g) {
      const int sl = c.k & 1;
      if (c.first_of_item()) mbar_wait(&bar_full[sl], (uint32_t)((c.k >> 1) & 1));
      const uint32_t q_b = smem_b + sl * slot_bytes, k_b = q_b + tb, v_b = k_b + tb, g_b = v_b + tb;
      const bool dq = c.is_dq();
      const int tile = c.tile();
      // lanes: dQ tile -> Q / dO rows of the tile against K / V;  dKV tile -> K / V rows against Q / dO
      const uint32_t a_s = (dq ? q_b : k_b) + tile * 16384, a_p = (dq ? g_b : v_b) + tile * 16384;
      const uint32_t b_s = dq ? k_b : q_b, b_p = dq ? v_b : g_b;
      const int c0 = c.gi * kGW, cw = min(kGW, npad - c0);
      const uint32_t Tb = tmem_base + (g & 1) * kBufCols;
      if (elect_one()) {
        const uint32_t idesc = idesc_bf16(128, cw, 0, 0);
        const uint64_t das = smem_desc_sw128(a_s), dap = smem_desc_sw128(a_p);
        const uint64_t dbs = smem_desc_sw128(b_s + c0 * 128), dbp = smem_desc_sw128(b_p + c0 * 128);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma_ss(Tb, das + 2 * kk, dbs + 2 * kk, idesc, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma_ss(Tb + kGW, dap + 2 * kk, dbp + 2 * kk, idesc, kk > 0);
        mma_commit(&bar_sfull[g & 1]);
      }
      __syncwarp();
    };
    for (int i = 0; i < 2; ++i) {
      if (cs.valid()) {
        issue_s(cs, gs++);
        cs.advance();
      }
    }
    while (ca.valid()) {
      const int sl = ca.k & 1;
      const uint32_t q_b = smem_b + sl * slot_bytes, k_b = q_b + tb, v_b = k_b + tb, g_b = v_b + tb;
      const bool dq = ca.is_dq();
      const int c0 = ca.gi * kGW, cw = min(kGW, npad - c0);
      const int h0 = ((cw >> 4) + 1) / 2 * 16, h1 = cw - h0;  // columns of the two warp halves
      const uint32_t Tb = tmem_base + (ga & 1) * kBufCols;
      mbar_wait(&bar_pfull[ga & 1], (ga >> 1) & 1);
      if (ca.gi == 0) mbar_wait(bar_accempty, (nt & 1) ^ 1);  // the previous lane tile's epilogue has read its accumulators
      tc_fence_after();
      if (lane == 0) dbg_stamp(p, 2, (int)ga * 2);
      if (elect_one()) {
        // A = P / dS (bf16) from TMEM, two runs (one per column half); B = the column-side operand, MN-major
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int hc0 = half == 0 ? 0 : h0, hw = half == 0 ? h0 : h1;
          const uint32_t adv = (uint32_t)(128 * ((c0 + hc0) >> 4));
          const int ksteps = hw >> 4;
          if (dq) {
            const uint64_t dk = smem_desc_sw128(k_b) + adv;
            for (int kk = 0; kk < ksteps; ++kk)  // dQ += dS K
              mma_ts(tmem_base + kAccA, Tb + hc0 + kk * 8, dk + 128 * kk, idesc_acc, (ca.gi | half | kk) != 0);
          } else {
            const uint64_t dg = smem_desc_sw128(g_b) + adv, dqd = smem_desc_sw128(q_b) + adv;
            for (int kk = 0; kk < ksteps; ++kk)  // dV += P^T dO
              mma_ts(tmem_base + kAccA, Tb + hc0 + kk * 8, dg + 128 * kk, idesc_acc, (ca.gi | half | kk) != 0);
            for (int kk = 0; kk < ksteps; ++kk)  // dK += dS^T Q
              mma_ts(tmem_base + kAccB, Tb + kGW + hc0 + kk * 8, dqd + 128 * kk, idesc_acc, (ca.gi | half | kk) != 0);
          }
        }
        if (ca.last_of_tile()) mma_commit(bar_accfull);
        if (ca.last_of_item()) mma_commit(&bar_empty[sl]);
      }
      __syncwarp();
      if (ca.last_of_tile()) ++nt;
      ++ga;
      ca.advance();
      if (cs.valid()) {  // S / dP two groups ahead, into the buffer the accumulating MMAs above have just been queued on
        issue_s(cs, gs++);
        cs.advance();
      }
      if (lane == 0) dbg_stamp(p, 2, (int)(ga - 1) * 2 + 1);
    }
  } else if (warp == 10) {
    // ============================================ statistics helper =======================================
    // stats[k & 1][i] = (lse_i * log2 e, delta_i = sum_d dO[i][d] O[i][d]) for the rows of item k, one item ahead
    int k = 0;
    for (int bh = blockIdx.x; bh < total; bh += gridDim.x, ++k) {
      const int sl = k & 1;
      const int b = bh / H, h = bh - b * H;
      float2* st = stats + sl * 256;
      const uint8_t* sG = smem + sl * slot_bytes + 3 * tb;
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
    // ============================================ compute warps ===========================================
    const int ch = warp >> 2, wq = warp & 3;  // column half, TMEM lane quarter
    const int lrow = wq * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint8_t* my_stage = tail + kTailStage + warp * 2048;
    const bool leader = elect_one();  // issues, commits and waits for this warp's TMA stores
    const bool dbg_on = p.dbg != nullptr && (tid & 127) == 0;
    Cursor c{0, (int)blockIdx.x, 0, 0, (int)gridDim.x, total, ntile, ngr};
    uint32_t g = 0, nt = 0;
    float neg_lse2 = 0.f, delta = 0.f;
    bool pend = false, pend_dq = false;
    int pend_tile = 0, pend_bh = 0;
    auto epilogue = [&](bool e_dq, int e_tile, int e_bh) {
      mbar_wait(bar_accfull, nt & 1);
      tc_fence_after();
      const int row0 = e_tile * 128 + wq * 32;
      if (row0 < N) {
        if (e_dq) {
          uint32_t a0[32];
          tmem_ld32(trow + kAccA + ch * 32, a0);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(bar_accempty);
          stage_and_store(&tm_out, my_stage, leader, a0, p.scale, lane, ch * 32, row0, e_bh);
        } else {
          uint32_t a0[32], a1[32];
          tmem_ld32(trow + (ch == 0 ? kAccA : kAccB), a0);
          tmem_ld32(trow + (ch == 0 ? kAccA : kAccB) + 32, a1);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(bar_accempty);
          const int plane = (ch == 0 ? 2 : 1) * p.B * H + e_bh;  // half 0: dV, half 1: dK
          const float mul = ch == 0 ? 1.f : p.scale;
          stage_and_store(&tm_out, my_stage, leader, a0, mul, lane, 0, row0, plane);
          stage_and_store(&tm_out, my_stage, leader, a1, mul, lane, 32, row0, plane);
        }
      } else {
        tc_fence_before();
        mbar_arrive(bar_accempty);
      }
      ++nt;
    };
    while (c.valid()) {
      const int sl = c.k & 1;
      const float2* st = stats + sl * 256;
      if (c.first_of_item()) mbar_wait(&bar_stfull[sl], (uint32_t)((c.k >> 1) & 1));
      const bool dq = c.is_dq();
      const int tile = c.tile();
      const int idx = tile * 128 + lrow;  // query row (dQ tile) / key (dKV tile) of this thread
      const bool warp_live = tile * 128 + wq * 32 < N;
      if (c.gi == 0) {
        const float2 mine = st[min(idx, 255)];
        neg_lse2 = -mine.x;
        delta = mine.y;
      }
      const int c0 = c.gi * kGW, cw = min(kGW, npad - c0);
      const int h0 = ((cw >> 4) + 1) / 2 * 16;
      const int hc0 = ch == 0 ? 0 : h0, hw = ch == 0 ? h0 : cw - h0;  // this warp's columns of the group
      const uint32_t Tb = trow + (g & 1) * kBufCols;
      if (dbg_on) dbg_stamp(p, ch, (int)g * 4 + 0);
      mbar_wait(&bar_sfull[g & 1], (g >> 1) & 1);
      tc_fence_after();
      if (dbg_on) dbg_stamp(p, ch, (int)g * 4 + 1);
      if (warp_live && hw > 0) {
        const float4* st2 = reinterpret_cast<const float4*>(st + c0 + hc0);
        const uint32_t t_s = Tb + hc0, t_dp = Tb + kGW + hc0;
        const int n16 = hw >> 4;
        if (dq) {
#pragma unroll 1
          for (int q = 0; q < n16; ++q)
            process16<true>(p, t_s + q * 16, t_dp + q * 16, 0, t_s + q * 8, c0 + hc0 + q * 16, neg_lse2, delta, st2);
        } else {
#pragma unroll 1
          for (int q = 0; q < n16; ++q)
            process16<false>(p, t_s + q * 16, t_dp + q * 16, t_s + q * 8, t_dp + q * 8, 0, 0.f, 0.f, st2 + q * 8);
        }
        tmem_wait_st();
      }
      tc_fence_before();
      mbar_arrive(&bar_pfull[g & 1]);
      if (dbg_on) dbg_stamp(p, ch, (int)g * 4 + 2);
      // ---- epilogue of the PREVIOUS lane tile, one group late: its accumulating MMAs ran while this group was in its
      // exponentials, so the wait below is (almost) free; warps of column half 0 take dQ channels 0..31 / dV, half 1
      // dQ channels 32..63 / dK
      if (pend) {
        epilogue(pend_dq, pend_tile, pend_bh);
        pend = false;
      }
      if (c.last_of_tile()) {
        pend = true;
        pend_dq = dq;
        pend_tile = tile;
        pend_bh = c.bh;
      }
      ++g;
      c.advance();
    }
    if (pend) epilogue(pend_dq, pend_tile, pend_bh);
    if (leader) bulk_wait<0>();  // stores complete before the CTA (and its shared memory) goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, kTmemCols);
}

int tensor_bytes(int N) { return ((((N + 7) & ~7) * 128) + 1023) & ~1023; }
size_t bws2_smem_bytes(int N) { return 1024 + (size_t)8 * tensor_bytes(N) + kTailBytes; }

std::atomic<long long*> g_bws2_dbg{nullptr};

}  // namespace

void attn_bwd_ws2_set_debug(long long* buf) { g_bws2_dbg.store(buf); }

bool attn_bwd_ws2_supported(int B, int H, int N, int Dh, const vrr_bias_desc* bias) {
  if (Dh != kDh || N < 1 || N > 256) return false;
  if (bias && bias->mode != VRR_BIAS_NONE) return false;
  if ((long long)3 * B * H * N >= (1ll << 31)) return false;
  static_assert(kTailStage + 8 * 2048 <= kTailBytes && kTailBytes >= 16384 && kTailStage % 128 == 0, "tail layout");
  return bws2_smem_bytes(N) <= (size_t)kSmemMax;
}

int attn_bwd_ws2(const void* planes, const void* out, const void* d_out, const float* lse, void* d_planes, int B, int H,
                 int N, int Dh, float scale, cudaStream_t st) {
  (void)Dh;
  VRR_REQUIRE(((uintptr_t)planes & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)d_out & 15) == 0 &&
                  ((uintptr_t)d_planes & 15) == 0,
              VRR_ERR_INVALID_ARG, "attn_bwd (tcgen05): planes / out / d_out / d_planes must be 16-byte aligned");
  const int E = H * kDh;
  const int rows8 = (N + 7) & ~7, trows = rows8 & 63;
  CUtensorMap tm_pl64, tm_plt, tm_do64, tm_dot, tm_out;
  if (int rc = make_tmap_2d(&tm_pl64, planes, 2, (uint64_t)3 * B * H * N, kDh, kDh * 2, 64, 64)) return rc;
  if (int rc = make_tmap_2d(&tm_plt, planes, 2, (uint64_t)3 * B * H * N, kDh, kDh * 2, trows ? trows : 8, 64)) return rc;
  if (int rc = make_tmap_2d(&tm_do64, d_out, 2, (uint64_t)B * N, (uint64_t)E, (uint64_t)E * 2, 64, 64)) return rc;
  if (int rc = make_tmap_2d(&tm_dot, d_out, 2, (uint64_t)B * N, (uint64_t)E, (uint64_t)E * 2, trows ? trows : 8, 64)) return rc;
  // d_planes[3 B H][N][64], box = [1][32 rows][32 channels], no swizzle (plain [32][64 B] staging)
  if (int rc = make_tmap_3d_bf16(&tm_out, d_planes, 64, (uint64_t)N, (uint64_t)3 * B * H, 128, (uint64_t)N * 128, 32, 32, 0)) return rc;
  Bws2Params p;
  p.out = (const __nv_bfloat16*)out;
  p.lse = lse;
  p.B = B; p.H = H; p.N = N;
  p.scale = scale;
  p.scale_log2 = scale * kLog2e;
  p.total_items = B * H;
  p.tb = tensor_bytes(N);
  p.dbg = g_bws2_dbg.load();
  const size_t smem = bws2_smem_bytes(N);
  const int grid = p.total_items < sm_count() ? p.total_items : sm_count();
  VRR_SMEM_ATTR_ONCE(attn_bwd_ws2_kernel, kSmemMax);
  attn_bwd_ws2_kernel<<<grid, kThreads, smem, st>>>(tm_pl64, tm_plt, tm_do64, tm_dot, tm_out, p);
  VRR_LAUNCHED();
  return VRR_OK;
}

}  // namespace vrr
