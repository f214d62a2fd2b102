// QKV projection on tcgen05 tensor cores with the RoPE rotate-half epilogue (bf16, head dim 64).
//
// Replaces models/vit.py:47-68 + models/rope_utils.py:22-35 of the reference: qkv Linear, the
// reshape/permute into heads, the cls/patch split, the rotation of q and k and the re-concatenation,
// in ONE kernel: planes[3][B][H][N][64] = rope(split_heads(x . w_qkv^T)).
//
// Persistent, warp-specialised kernel, one CTA per SM, tile 128 x BN (BN = 256 or 192 output
// columns = 4 or 3 whole heads):
//   warp 0      TMA producer: x tile [128 x 64] and w tile [BN x 64] per k-block into a 4-stage ring
//               (cp.async.bulk.tensor, 128-byte swizzle, mbarrier complete_tx);
//   warp 1      MMA issuer: tcgen05.mma kind::f16 (M 128, N BN, K 16), fp32 accumulators in TMEM,
//               double-buffered (2 x BN columns) so the epilogue of tile i overlaps the mainloop of i+1;
//   warps 2-5   epilogue: tcgen05.ld one head (64 columns) of the thread's row, rotate the pairs
//               (d, d+32) with cos/sin of the row's patch position (cls row and V untouched), round to
//               bf16 and write the 128-byte head row straight into its [which][b][h][t][:] slot.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrr {

using namespace tc;

namespace {

constexpr int kBM = 128, kBK = 64, kGemmStages = 4;
constexpr int kATileBytes = kBM * kBK * 2;  // 16 KB

struct QkvParams {
  __nv_bfloat16* planes;
  const float *cos_tab, *sin_tab;
  int B, N, E, H, rope_mode;
  int M;        // B * N
  int tiles_m, tiles_n, num_k;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
qkv_rope_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const QkvParams p) {
  constexpr int kBTileBytes = BN * kBK * 2;
  constexpr int kStageBytes = kATileBytes + kBTileBytes;
  constexpr uint32_t kTmemCols = 512;  // 2 accumulator stages x BN columns, power of two
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kStageBytes);
  uint64_t* bar_full = bars;                       // [stages]  TMA -> MMA
  uint64_t* bar_empty = bars + kGemmStages;        // [stages]  MMA -> TMA
  uint64_t* bar_acc_full = bars + 2 * kGemmStages;       // [2]  MMA -> epilogue
  uint64_t* bar_acc_empty = bars + 2 * kGemmStages + 2;  // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGemmStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_acc_full[a], 1);
      mbar_init(&bar_acc_empty[a], 128);
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_x);
      tma_prefetch_desc(&tmap_w);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_expect_tx(&bar_full[stage], kStageBytes);
          tma_load_2d(sa, &tmap_x, &bar_full[stage], kb * kBK, m_blk * kBM);
          tma_load_2d(sa + kATileBytes, &tmap_w, &bar_full[stage], kb * kBK, n_blk * BN);
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(kBM, BN, 0, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&bar_acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(&bar_full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + kATileBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) mma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          mma_commit(&bar_empty[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&bar_acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int hd = 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      const int m = m_blk * kBM + row;
      const bool live = m < p.M;
      const int b = live ? m / p.N : 0, t = live ? m - b * p.N : 0;
      mbar_wait(&bar_acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int hc = 0; hc < BN / 64; ++hc) {
        uint32_t v[64];
        __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
        tmem_ld32(trow + hc * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(trow + hc * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_wait_ld();
        if (hc == BN / 64 - 1) {  // all TMEM reads of this accumulator are done: hand it back
          tc_fence_before();
          mbar_arrive(&bar_acc_empty[acc]);
        }
        if (live) {
        const int n0 = n_blk * BN + hc * 64;
        const int which = n0 / p.E, h = (n0 - which * p.E) >> 6;
        float* f = reinterpret_cast<float*>(v);
        if (p.rope_mode != VRR_ROPE_NONE && which < 2 && t >= 1) {
          const size_t base = ((size_t)(p.rope_mode == VRR_ROPE_MIXED ? h * (p.N - 1) : 0) + (t - 1)) * hd;
          const float4* c4 = reinterpret_cast<const float4*>(p.cos_tab + base);
          const float4* s4 = reinterpret_cast<const float4*>(p.sin_tab + base);
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
        __nv_bfloat16* dst = p.planes + ((((size_t)which * p.B + b) * p.H + h) * p.N + t) * 64;
#pragma unroll
        for (int v8 = 0; v8 < 8; ++v8) {
          uint4 w;
          w.x = pack_bf16(f[v8 * 8 + 0], f[v8 * 8 + 1]);
          w.y = pack_bf16(f[v8 * 8 + 2], f[v8 * 8 + 3]);
          w.z = pack_bf16(f[v8 * 8 + 4], f[v8 * 8 + 5]);
          w.w = pack_bf16(f[v8 * 8 + 6], f[v8 * 8 + 7]);
          *reinterpret_cast<uint4*>(dst + v8 * 8) = w;
        }
        }  // live
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN>
int launch_qkv(const void* x, const void* w, const float* cos_tab, const float* sin_tab, void* planes, int B, int N,
               int E, int H, int rope_mode, cudaStream_t st) {
  CUtensorMap tx, tw;
  const int M = B * N;
  if (int rc = make_tmap_bf16(&tx, x, (uint64_t)M, (uint64_t)E, (uint64_t)E * 2, kBM)) return rc;
  if (int rc = make_tmap_bf16(&tw, w, (uint64_t)3 * E, (uint64_t)E, (uint64_t)E * 2, BN)) return rc;
  QkvParams p;
  p.planes = (__nv_bfloat16*)planes;
  p.cos_tab = cos_tab; p.sin_tab = sin_tab;
  p.B = B; p.N = N; p.E = E; p.H = H; p.rope_mode = rope_mode; p.M = M;
  p.tiles_m = ceil_div(M, kBM);
  p.tiles_n = 3 * E / BN;
  p.num_k = E / kBK;
  const size_t smem = 1024 + (size_t)kGemmStages * (kATileBytes + BN * kBK * 2) + (2 * kGemmStages + 4) * 8 + 16;
  static bool attr_set = false;
  if (!attr_set) {
    VRR_CUDA(cudaFuncSetAttribute(qkv_rope_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = min(sm_count(), p.tiles_m * p.tiles_n);
  qkv_rope_tc_kernel<BN><<<grid, 192, smem, st>>>(tx, tw, p);
  VRR_LAUNCHED();
  return VRR_OK;
}

}  // namespace

bool qkv_rope_fwd_tc_supported(int B, int N, int E, int H) {
  if (H <= 0 || E % H != 0 || E / H != 64) return false;
  if ((long long)B * N >= (1ll << 31) / 4) return false;
  return true;
}

int qkv_rope_fwd_tc(const void* x, const void* w, const float* cos_tab, const float* sin_tab, void* planes, int B,
                    int N, int E, int H, int rope_mode, cudaStream_t st) {
  VRR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)planes & 15) == 0,
              VRR_ERR_INVALID_ARG, "qkv_rope_fwd (tcgen05): x / w / planes must be 16-byte aligned");
  if (rope_mode != VRR_ROPE_NONE)
    VRR_REQUIRE(((uintptr_t)cos_tab & 15) == 0 && ((uintptr_t)sin_tab & 15) == 0, VRR_ERR_INVALID_ARG,
                "qkv_rope_fwd (tcgen05): cos / sin must be 16-byte aligned");
  if ((3 * E) % 256 == 0) return launch_qkv<256>(x, w, cos_tab, sin_tab, planes, B, N, E, H, rope_mode, st);
  return launch_qkv<192>(x, w, cos_tab, sin_tab, planes, B, N, E, H, rope_mode, st);
}

}  // namespace vrr
