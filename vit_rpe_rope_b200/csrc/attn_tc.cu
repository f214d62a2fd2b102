// Fused attention forward on the 5th-generation tensor cores (tcgen05 + TMEM), bf16, head dim 64.
//
// Replaces models/vit.py:71-88 of the reference: S = q k^T * scale (+ relative-position table /
// polynomial bias) -> softmax -> . v -> merge heads, without materialising S or P in HBM.
//
// One CTA = 128 query rows of one (image, head); 128 threads, thread t owns query row t = TMEM lane t.
//   * Q tile and 128-key K / V tiles arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a
//     2-stage ring; for the 65..197-token ViT sequences the whole sequence is resident in shared
//     memory after the prologue, longer sequences (577, 1025 tokens) stream through the ring.
//   * S = Q K^T : tcgen05.mma kind::f16, both operands from shared memory, fp32 accumulator in
//     TMEM columns [0,128).
//   * softmax  : each thread pulls its row with tcgen05.ld, applies scale*log2e and the bias
//     (table row staged into shared memory by a TMA bulk copy / polynomial evaluated into a
//     distance LUT), keeps running max / sum (online softmax across key tiles), packs P to bf16
//     and writes it back over S with tcgen05.st (columns [0,64)).
//   * O_tile = P V : tcgen05.mma with A = P from TMEM and B = V from shared memory (MN-major),
//     accumulator in TMEM columns [128,192); the thread folds it into its fp32 register accumulator
//     with the online-softmax correction.
//   * epilogue: O / l -> bf16 -> out[b][i][h*64 ..], lse.
// 256 TMEM columns and ~85 KB shared memory per CTA -> two CTAs per SM, so one CTA's softmax
// overlaps the other's TMA / MMA.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrr {

using namespace tc;

namespace {

constexpr int kM = 128;       // query rows per CTA
constexpr int kKT = 128;      // keys per tile
constexpr int kStages = 2;    // K/V ring depth
constexpr int kDh = 64;
constexpr int kTileBytes = kKT * kDh * 2;  // 16 KB: one 128-row x 64-col bf16 tile
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColO = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct FwdParams {
  __nv_bfloat16* out;
  float* lse;
  const float* bias_param;
  int B, H, N;
  float scale_log2;  // scale * log2(e)
  int bias_heads, bias_len, bias_grid;
  int lut_floats;    // shared-memory floats reserved for the LUT (incl. alignment slack)
};

template <int BIAS>
__global__ void __launch_bounds__(128, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: SWIZZLE_128B atoms and UMMA descriptors with base_offset = 0
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;                 // [kStages][16 KB]
  uint8_t* sV = sK + kStages * kTileBytes;       // [kStages][16 KB]
  float* lut_raw = reinterpret_cast<float*>(sV + kStages * kTileBytes);
  uint16_t* key_yx = reinterpret_cast<uint16_t*>(lut_raw + p.lut_floats);  // POLY: (y << 8) | x per token
  const int key_yx_bytes = (BIAS == VRR_BIAS_POLY) ? ((p.N * 2 + 15) & ~15) : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(key_yx) + key_yx_bytes);
  uint64_t* bar_full = bars;                 // [kStages]
  uint64_t* bar_q = bars + kStages;
  uint64_t* bar_s = bars + kStages + 1;
  uint64_t* bar_o = bars + kStages + 2;
  uint64_t* bar_lut = bars + kStages + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = p.N, H = p.H;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int m0 = blockIdx.x * kM;
  const int i = m0 + tid;
  const bool warp_active = (m0 + warp * 32) < N;
  const int ntiles = (N + kKT - 1) / kKT;
  const int BHN = p.B * H * N;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bar_full[s], 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_lut, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);  // this warp's lane quadrant

  // ---- bias LUT (values pre-multiplied by log2 e) -------------------------------------------------
  const float* lut = lut_raw;
  if (BIAS == VRR_BIAS_TABLE) {
    // TMA bulk copy of the head's table row.  Rows are 2N-1 floats, i.e. not 16-byte aligned in
    // general: copy the enclosing aligned span that stays inside the table, finish the (<= 3
    // element) tail with plain loads.
    const size_t row_b = (size_t)h * p.bias_len * 4, row_e = row_b + (size_t)p.bias_len * 4;
    const size_t tot = (size_t)p.bias_heads * p.bias_len * 4;
    const size_t a0 = row_b & ~size_t(15);
    size_t a1 = (row_e + 15) & ~size_t(15);
    if (a1 > (tot & ~size_t(15))) a1 = tot & ~size_t(15);
    const uint32_t bulk = a1 > a0 ? (uint32_t)(a1 - a0) : 0u;
    lut = lut_raw + (row_b - a0) / 4;
    if (tid == 0) {
      if (bulk) {
        mbar_expect_tx(bar_lut, bulk);
        bulk_load_1d(lut_raw, reinterpret_cast<const uint8_t*>(p.bias_param) + a0, bulk, bar_lut);
      } else {
        mbar_arrive(bar_lut);
      }
    }
  }

  // ---- TMA prologue: Q tile + the first kStages K/V tiles ----------------------------------------
  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    mbar_expect_tx(bar_q, kTileBytes);
    tma_load_2d(sQ, &tmap, bar_q, 0, bh * N + m0);
    for (int s = 0; s < kStages && s < ntiles; ++s) {
      mbar_expect_tx(&bar_full[s], 2 * kTileBytes);
      tma_load_2d(sK + s * kTileBytes, &tmap, &bar_full[s], 0, BHN + bh * N + s * kKT);
      tma_load_2d(sV + s * kTileBytes, &tmap, &bar_full[s], 0, 2 * BHN + bh * N + s * kKT);
    }
  }
  __syncwarp();

  int yi = 0, xi = 0;
  if (BIAS == VRR_BIAS_TABLE) {
    mbar_wait(bar_lut, 0);
    const size_t row_b = (size_t)h * p.bias_len * 4, row_e = row_b + (size_t)p.bias_len * 4;
    const size_t tot = (size_t)p.bias_heads * p.bias_len * 4;
    size_t a1 = (row_e + 15) & ~size_t(15);
    if (a1 > (tot & ~size_t(15))) a1 = tot & ~size_t(15);
    const int have = a1 > row_b ? (int)((a1 - row_b) / 4) : 0;  // row elements delivered by the bulk copy
    float* lw = const_cast<float*>(lut);
    const float* grow = p.bias_param + (size_t)h * p.bias_len;
    for (int t = tid; t < p.bias_len; t += 128) lw[t] = (t < have ? lw[t] : grow[t]) * kLog2e;
    __syncthreads();
  } else if (BIAS == VRR_BIAS_POLY) {
    float* lw = lut_raw;
    const float* c = p.bias_param + (size_t)(p.bias_heads == 1 ? 0 : h) * p.bias_len;
    for (int d = tid; d < 2 * p.bias_grid - 1; d += 128) {
      float x = (float)d, pw = 1.f, acc = 0.f;
      for (int k = 0; k < p.bias_len; ++k) {
        acc = fmaf(pw, c[k], acc);
        pw *= x;
      }
      lw[d] = acc * kLog2e;
    }
    for (int t = tid; t < N; t += 128) {
      int pt = t > 0 ? t - 1 : 0;
      key_yx[t] = (uint16_t)(((pt % p.bias_grid) << 8) | (pt / p.bias_grid));
    }
    const int pi = i > 0 ? i - 1 : 0;
    yi = pi % p.bias_grid;
    xi = pi / p.bias_grid;
    __syncthreads();
  }

  const uint64_t desc_q = smem_desc_sw128(smem_u32(sQ));
  constexpr uint32_t idesc_o = idesc_bf16(kM, kDh, 0, 1);  // P (TMEM) x V (MN-major shared)

  float acc[kDh];
#pragma unroll
  for (int d = 0; d < kDh; ++d) acc[d] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int t = 0; t < ntiles; ++t) {
    const int stage = t % kStages;
    const uint32_t use_parity = (uint32_t)((t / kStages) & 1);
    const int nvalid = min(kKT, N - t * kKT);
    const int npad = (nvalid + 15) & ~15;

    // ---- S = Q K^T ---------------------------------------------------------------------------
    if (tid == 0) {
      if (t == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_full[stage], use_parity);
      tc_fence_after();
      const uint64_t desc_k = smem_desc_sw128(smem_u32(sK + stage * kTileBytes));
      const uint32_t idesc_s = idesc_bf16(kM, npad, 0, 0);
#pragma unroll
      for (int k = 0; k < kDh / 16; ++k)  // 16 bf16 = 32 bytes along the swizzled row: +2 encoded units
        mma_ss(tmem_base + kColS, desc_q + 2 * k, desc_k + 2 * k, idesc_s, k > 0);
      mma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, (uint32_t)(t & 1));
    tc_fence_after();

    // ---- softmax on this thread's row -----------------------------------------------------------
    float corr = 1.f;
    if (warp_active) {
      const int nchunks = (nvalid + 31) >> 5;
      uint32_t sraw[kKT];
#pragma unroll
      for (int c = 0; c < kKT / 32; ++c)
        if (c < nchunks) tmem_ld32(tmem_row + kColS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sraw[c * 32]));
      tmem_wait_ld();
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < kKT / 32; ++c) {
        if (c < nchunks) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int jl = c * 32 + e, j = t * kKT + jl;
            float v = __uint_as_float(sraw[jl]) * p.scale_log2;
            if (BIAS == VRR_BIAS_TABLE) {
              v += lut[min(max(i - j + N - 1, 0), 2 * N - 2)];
            } else if (BIAS == VRR_BIAS_POLY) {
              const int yx = key_yx[min(j, N - 1)];
              const int dist = abs(yi - (yx >> 8)) + abs(xi - (yx & 255));
              v += (i == 0 || j == 0) ? 0.f : lut[dist];
            }
            v = (jl < nvalid) ? v : -INFINITY;
            sraw[jl] = __float_as_uint(v);
            tmax = fmaxf(tmax, v);
          }
        }
      }
      const float m_new = fmaxf(m_run, tmax);
      corr = ex2(m_run - m_new);
      l_run *= corr;
      m_run = m_new;
#pragma unroll
      for (int c = 0; c < kKT / 32; ++c) {
        if (c < nchunks) {
          uint32_t packed[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = ex2(__uint_as_float(sraw[c * 32 + 2 * e]) - m_new);
            const float p1 = ex2(__uint_as_float(sraw[c * 32 + 2 * e + 1]) - m_new);
            l_run += p0 + p1;
            packed[e] = pack_bf16(p0, p1);
          }
          tmem_st16(tmem_row + kColS + c * 16, packed);  // P (bf16 pairs) overwrites S in place
        }
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();

    // ---- O_tile = P V --------------------------------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
      const uint64_t desc_v = smem_desc_sw128(smem_u32(sV + stage * kTileBytes));
      const int ksteps = npad >> 4;
      for (int kk = 0; kk < ksteps; ++kk)  // 16 keys = 16 rows x 128 B = 2048 B: +128 encoded units
        mma_ts(tmem_base + kColO, tmem_base + kColS + kk * 8, desc_v + 128 * kk, idesc_o, kk > 0);
      mma_commit(bar_o);
    }
    __syncwarp();
    mbar_wait(bar_o, (uint32_t)(t & 1));
    tc_fence_after();

    // this stage's K/V are consumed: refill it with tile t + kStages
    if (tid == 0 && t + kStages < ntiles) {
      const int tn = t + kStages;
      mbar_expect_tx(&bar_full[stage], 2 * kTileBytes);
      tma_load_2d(sK + stage * kTileBytes, &tmap, &bar_full[stage], 0, BHN + bh * N + tn * kKT);
      tma_load_2d(sV + stage * kTileBytes, &tmap, &bar_full[stage], 0, 2 * BHN + bh * N + tn * kKT);
    }
    __syncwarp();

    if (warp_active) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[32];
        tmem_ld32(tmem_row + kColO + half * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int d = 0; d < 32; ++d) acc[half * 32 + d] = fmaf(acc[half * 32 + d], corr, __uint_as_float(o[d]));
      }
    }
  }

  // ---- epilogue -----------------------------------------------------------------------------------
  if (i < N) {
    const float inv = 1.f / l_run;
    __nv_bfloat16* dst = p.out + ((size_t)b * N + i) * (size_t)(H * kDh) + h * kDh;
#pragma unroll
    for (int v8 = 0; v8 < kDh / 8; ++v8) {
      uint4 w;
      w.x = pack_bf16(acc[v8 * 8 + 0] * inv, acc[v8 * 8 + 1] * inv);
      w.y = pack_bf16(acc[v8 * 8 + 2] * inv, acc[v8 * 8 + 3] * inv);
      w.z = pack_bf16(acc[v8 * 8 + 4] * inv, acc[v8 * 8 + 5] * inv);
      w.w = pack_bf16(acc[v8 * 8 + 6] * inv, acc[v8 * 8 + 7] * inv);
      *reinterpret_cast<uint4*>(dst + v8 * 8) = w;
    }
    p.lse[(size_t)bh * N + i] = (m_run + log2f(l_run)) * kLn2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

size_t fwd_smem_bytes(int N, const vrr_bias_desc* bias, int* lut_floats) {
  int lf = 0;
  if (bias && bias->mode == VRR_BIAS_TABLE) lf = 2 * N - 1 + 8;        // + slack for the aligned span
  else if (bias && bias->mode == VRR_BIAS_POLY) lf = 2 * bias->grid - 1;
  lf = (lf + 3) & ~3;
  if (lut_floats) *lut_floats = lf;
  size_t yx = (bias && bias->mode == VRR_BIAS_POLY) ? (size_t)((N * 2 + 15) & ~15) : 0;
  return 1024 + (size_t)kTileBytes * (1 + 2 * kStages) + (size_t)lf * 4 + yx + (kStages + 4) * 8 + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

}  // namespace

int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                   uint32_t box_rows) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VRR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    VRR_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, VRR_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VRR_REQUIRE(r == CUDA_SUCCESS, VRR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VRR_OK;
}

bool attn_fwd_tc_supported(int B, int H, int N, int Dh, const vrr_bias_desc* bias) {
  if (Dh != kDh || N < 1) return false;
  if ((long long)3 * B * H * N >= (1ll << 31)) return false;
  if (bias && bias->mode == VRR_BIAS_POLY && bias->grid > 255) return false;
  return fwd_smem_bytes(N, bias, nullptr) <= 110 * 1024;
}

int attn_fwd_tc(const void* planes, const vrr_bias_desc* bias, void* out, float* lse, int B, int H, int N, int Dh,
                float scale, cudaStream_t st) {
  (void)Dh;
  VRR_REQUIRE(((uintptr_t)planes & 15) == 0 && ((uintptr_t)out & 15) == 0, VRR_ERR_INVALID_ARG,
              "attn_fwd (tcgen05): planes/out must be 16-byte aligned");
  CUtensorMap tmap;
  if (int rc = make_tmap_bf16(&tmap, planes, (uint64_t)3 * B * H * N, kDh, kDh * 2, kKT)) return rc;
  FwdParams p;
  p.out = (__nv_bfloat16*)out;
  p.lse = lse;
  p.B = B; p.H = H; p.N = N;
  p.scale_log2 = scale * kLog2e;
  const int mode = bias ? bias->mode : VRR_BIAS_NONE;
  p.bias_param = mode != VRR_BIAS_NONE ? bias->param : nullptr;
  p.bias_heads = mode != VRR_BIAS_NONE ? bias->heads : 0;
  p.bias_len = mode != VRR_BIAS_NONE ? bias->len : 0;
  p.bias_grid = mode != VRR_BIAS_NONE ? bias->grid : 0;
  if (mode == VRR_BIAS_TABLE)
    VRR_REQUIRE(((uintptr_t)bias->param & 15) == 0, VRR_ERR_INVALID_ARG, "attn_fwd (tcgen05): bias table must be 16-byte aligned");
  const size_t smem = fwd_smem_bytes(N, bias, &p.lut_floats);
  dim3 grid(ceil_div(N, kM), B * H);
#define LAUNCH(MODE)                                                                                         \
  do {                                                                                                       \
    static bool attr_set = false;                                                                            \
    if (!attr_set) {                                                                                         \
      VRR_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                    110 * 1024));                                                            \
      attr_set = true;                                                                                       \
    }                                                                                                        \
    attn_fwd_tc_kernel<MODE><<<grid, 128, smem, st>>>(tmap, p);                                              \
  } while (0)
  if (mode == VRR_BIAS_TABLE) LAUNCH(VRR_BIAS_TABLE);
  else if (mode == VRR_BIAS_POLY) LAUNCH(VRR_BIAS_POLY);
  else LAUNCH(VRR_BIAS_NONE);
#undef LAUNCH
  VRR_LAUNCHED();
  return VRR_OK;
}

}  // namespace vrr
