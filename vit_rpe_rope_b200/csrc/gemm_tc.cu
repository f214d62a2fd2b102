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

struct GemmShape {
  int M;  // valid rows of A / C
  int tiles_m, tiles_n, num_k;
};

// ---- epilogues: one call = one 64-column chunk [n0, n0+64) of output row m, fp32 accumulators in f ----
struct QkvRopeEpi {
  __nv_bfloat16* planes;
  const float *cos_tab, *sin_tab;
  int B, N, E, H, rope_mode;
  __device__ __forceinline__ void operator()(int m, int n0, float (&f)[64]) const {
    const int hd = 32;
    const int b = m / N, t = m - b * N;
    const int which = n0 / E, h = (n0 - which * E) >> 6;
    if (rope_mode != VRR_ROPE_NONE && which < 2 && t >= 1) {
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
    __nv_bfloat16* dst = planes + ((((size_t)which * B + b) * H + h) * N + t) * 64;
#pragma unroll
    for (int v8 = 0; v8 < 8; ++v8) {
      uint4 w;
      w.x = pack_bf16(f[v8 * 8 + 0], f[v8 * 8 + 1]);
      w.y = pack_bf16(f[v8 * 8 + 2], f[v8 * 8 + 3]);
      w.z = pack_bf16(f[v8 * 8 + 4], f[v8 * 8 + 5]);
      w.w = pack_bf16(f[v8 * 8 + 6], f[v8 * 8 + 7]);
      *reinterpret_cast<uint4*>(dst + v8 * 8) = w;
    }
  }
};

// tokens[b][1+p][n] = round_bf16(acc + bias[n]) (+ pos[p][n]); TT = token-stream type (see gemm_simt.cu)
template <typename TT>
struct PatchEmbedEpi {
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

// C[m][n] = sum_k A[m][k] * Bw[n][k]  (both operands K-major bf16), 128 x BN tiles, persistent.
template <int BN, class Epi>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmShape g, const Epi epi) {
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
  const int total_tiles = g.tiles_m * g.tiles_n;

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
      tma_prefetch_desc(&tmap_a);
      tma_prefetch_desc(&tmap_b);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / g.tiles_n, n_blk = tile - m_blk * g.tiles_n;
        for (int kb = 0; kb < g.num_k; ++kb) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_expect_tx(&bar_full[stage], kStageBytes);
          tma_load_2d(sa, &tmap_a, &bar_full[stage], kb * kBK, m_blk * kBM);
          tma_load_2d(sa + kATileBytes, &tmap_b, &bar_full[stage], kb * kBK, n_blk * BN);
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
        for (int kb = 0; kb < g.num_k; ++kb) {
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
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / g.tiles_n, n_blk = tile - m_blk * g.tiles_n;
      const int m = m_blk * kBM + row;
      const bool live = m < g.M;
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
        if (live) epi(m, n_blk * BN + hc * 64, *reinterpret_cast<float(*)[64]>(v));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// A [M][K] and Bw [Nout][K], both bf16 row-major (K contiguous); Nout % BN == 0, K % 64 == 0.
template <int BN, class Epi>
int launch_gemm_tc(const void* a, const void* bw, int M, int Nout, int K, const Epi& epi, cudaStream_t st) {
  CUtensorMap ta, tb;
  if (int rc = make_tmap_bf16(&ta, a, (uint64_t)M, (uint64_t)K, (uint64_t)K * 2, kBM)) return rc;
  if (int rc = make_tmap_bf16(&tb, bw, (uint64_t)Nout, (uint64_t)K, (uint64_t)K * 2, BN)) return rc;
  GemmShape g;
  g.M = M;
  g.tiles_m = ceil_div(M, kBM);
  g.tiles_n = Nout / BN;
  g.num_k = K / kBK;
  const size_t smem = 1024 + (size_t)kGemmStages * (kATileBytes + BN * kBK * 2) + (2 * kGemmStages + 4) * 8 + 16;
  VRR_SMEM_ATTR_ONCE((gemm_tc_kernel<BN, Epi>), smem);
  const int grid = min(sm_count(), g.tiles_m * g.tiles_n);
  gemm_tc_kernel<BN, Epi><<<grid, 192, smem, st>>>(ta, tb, g, epi);
  VRR_LAUNCHED();
  return VRR_OK;
}

template <class Epi>
int launch_gemm_tc_any(const void* a, const void* bw, int M, int Nout, int K, const Epi& epi, cudaStream_t st) {
  if (Nout % 256 == 0) return launch_gemm_tc<256, Epi>(a, bw, M, Nout, K, epi, st);
  if (Nout % 192 == 0) return launch_gemm_tc<192, Epi>(a, bw, M, Nout, K, epi, st);
  if (Nout % 128 == 0) return launch_gemm_tc<128, Epi>(a, bw, M, Nout, K, epi, st);
  return launch_gemm_tc<64, Epi>(a, bw, M, Nout, K, epi, st);
}

// unfold(images)[m][k], m = b*Np + py*gw + px, k = c*P*P + i*P + j  ->  bf16; 8 consecutive j per thread
template <typename TI>
__global__ void patch_unfold_kernel(const TI* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int C, int Hi,
                                    int Wi, int P, int gw, int Np) {
  const int K8 = C * P * P / 8;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * Np * K8) return;
  const int k8 = (int)(idx % K8);
  const size_t m = idx / K8;
  const int b = (int)(m / Np), p = (int)(m - (size_t)b * Np), py = p / gw, px = p - py * gw;
  const int k = k8 * 8, c = k / (P * P), r = k - c * P * P, i = r / P, j = r - i * P;
  const TI* src = img + (((size_t)b * C + c) * Hi + (py * P + i)) * Wi + (px * P + j);
  float4 lo = ld4(src), hi = ld4(src + 4);
  __nv_bfloat16* dst = out + m * (size_t)(K8 * 8) + k;
  st4(dst, lo);
  st4(dst + 4, hi);
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
  QkvRopeEpi epi{(__nv_bfloat16*)planes, cos_tab, sin_tab, B, N, E, H, rope_mode};
  // whole heads per tile: 3E = 192 * H, so 256 | 3E or 192 | 3E always holds
  if ((3 * E) % 256 == 0) return launch_gemm_tc<256, QkvRopeEpi>(x, w, B * N, 3 * E, E, epi, st);
  return launch_gemm_tc<192, QkvRopeEpi>(x, w, B * N, 3 * E, E, epi, st);
}

// ---- patch embedding on the tensor cores: unfold pre-pass (fused fp32->bf16 cast) + GEMM ----------
bool patch_embed_tc_supported(int B, int C, int Hi, int Wi, int P, int E) {
  const int K = C * P * P;
  if (P % 8 != 0 || Wi % 4 != 0 || K % 64 != 0 || E % 64 != 0) return false;
  if ((long long)B * (Hi / P) * (Wi / P) >= (1ll << 31) / 4) return false;
  return true;
}
size_t patch_embed_tc_workspace_bytes(int B, int C, int Hi, int Wi, int P) {
  return (size_t)B * (Hi / P) * (Wi / P) * C * P * P * 2;
}
int patch_unfold(const void* images, void* out, int B, int C, int Hi, int Wi, int P, int img_dtype, cudaStream_t st) {
  const int gw = Wi / P, Np = (Hi / P) * gw;
  VRR_REQUIRE(P % 8 == 0 && Wi % 4 == 0, VRR_ERR_UNSUPPORTED, "patch_unfold: needs P %% 8 == 0 and Wi %% 4 == 0");
  VRR_REQUIRE(((uintptr_t)images & 15) == 0 && ((uintptr_t)out & 15) == 0, VRR_ERR_INVALID_ARG,
              "patch_unfold: pointers must be 16-byte aligned");
  const size_t total = (size_t)B * Np * (C * P * P / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (img_dtype == VRR_F32)
    patch_unfold_kernel<float><<<blocks, 256, 0, st>>>((const float*)images, (__nv_bfloat16*)out, B, C, Hi, Wi, P, gw, Np);
  else
    patch_unfold_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)images, (__nv_bfloat16*)out, B, C,
                                                              Hi, Wi, P, gw, Np);
  VRR_LAUNCHED();
  return VRR_OK;
}
int patch_embed_fwd_tc(const void* images, const void* weight, const void* bias, const void* pos, void* tokens,
                       void* workspace, int B, int C, int Hi, int Wi, int P, int E, int img_dtype, int tok_dtype,
                       cudaStream_t st) {
  const int Np = (Hi / P) * (Wi / P), K = C * P * P;
  if (int rc = patch_unfold(images, workspace, B, C, Hi, Wi, P, img_dtype, st)) return rc;
  if (tok_dtype == VRR_F32) {
    PatchEmbedEpi<float> epi{(float*)tokens, (const __nv_bfloat16*)bias, (const float*)pos, Np, E};
    return launch_gemm_tc_any(workspace, weight, B * Np, E, K, epi, st);
  }
  PatchEmbedEpi<__nv_bfloat16> epi{(__nv_bfloat16*)tokens, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)pos, Np, E};
  return launch_gemm_tc_any(workspace, weight, B * Np, E, K, epi, st);
}

}  // namespace vrr
