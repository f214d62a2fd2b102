// RoPE rotate-half kernels that are not GEMM epilogues:
//   * backward of the fused QKV+RoPE epilogue: un-rotate the plane gradients into the token
//     layout [B*N][3E] the two projection-backward GEMMs consume, and reduce d_cos / d_sin over
//     the batch (autograd of models/rope_utils.py:22-35 and of the cat/split at vit.py:56-68);
//   * the stand-alone public apply_rotary_emb (models/rope_utils.py:3-37).
#include "common.cuh"

namespace vrr {

// One thread = VEC consecutive channels dd..dd+VEC-1 of the first half (and their partners dd+Dh/2..)
// of one (head, token), looping over a chunk of the batch: 16-byte loads/stores throughout.
template <typename T, int VEC>
struct VecIO;
template <>
struct VecIO<float, 4> {
  using Raw = float4;
  static __device__ __forceinline__ void cvt(const float4& x, float (&v)[4]) {
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  }
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    cvt(*reinterpret_cast<const float4*>(p), v);
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct VecIO<__nv_bfloat16, 8> {
  using Raw = uint4;
  static __device__ __forceinline__ void cvt(const uint4& x, float (&v)[8]) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
      v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    cvt(*reinterpret_cast<const uint4*>(p), v);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 x;
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    x.x = *reinterpret_cast<uint32_t*>(&a); x.y = *reinterpret_cast<uint32_t*>(&b);
    x.z = *reinterpret_cast<uint32_t*>(&c); x.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = x;
  }
};

template <typename T, int VEC>
__global__ void qkv_rope_bwd_kernel(const T* __restrict__ d_planes, const T* __restrict__ planes,
                                    const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                    T* __restrict__ d_qkv, float* d_cos, float* d_sin, int B, int N,
                                    int E, int H, int Dh, int rope_mode, int b_chunk) {
  const int hd = Dh >> 1, nv = hd / VEC;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * N * nv) return;
  const int dd = (idx % nv) * VEC, t = (idx / nv) % N, h = idx / (nv * N);
  const int b0 = blockIdx.y * b_chunk, b1 = min(B, b0 + b_chunk);
  const size_t plane = (size_t)B * H * N * Dh;
  const bool rot = rope_mode != VRR_ROPE_NONE && t >= 1;
  float c[VEC], s[VEC], acc_c[VEC], acc_s[VEC];
  size_t tab = 0;
#pragma unroll
  for (int e = 0; e < VEC; ++e) { c[e] = 1.f; s[e] = 0.f; acc_c[e] = 0.f; acc_s[e] = 0.f; }
  if (rot) {
    tab = ((size_t)(rope_mode == VRR_ROPE_MIXED ? h * (N - 1) : 0) + (t - 1)) * hd + dd;
#pragma unroll
    for (int e = 0; e < VEC; ++e) { c[e] = cos_tab[tab + e]; s[e] = sin_tab[tab + e]; }
  }
  const bool want_cs = rot && d_cos != nullptr;
  // All ten 16-byte loads of one image are issued before the first use (raw, converted when consumed): with the
  // loads interleaved with the stores a thread had 32 bytes in flight and the kernel ran at a third of the HBM rate.
  using Raw = typename VecIO<T, VEC>::Raw;
  for (int b = b0; b < b1; ++b) {
    const size_t src = (((size_t)b * H + h) * N + t) * Dh + dd;
    T* dst = d_qkv + ((size_t)b * N + t) * (3 * E) + h * Dh + dd;
    Raw rg[3][2], rx[2][2];
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      rg[which][0] = *reinterpret_cast<const Raw*>(d_planes + which * plane + src);
      rg[which][1] = *reinterpret_cast<const Raw*>(d_planes + which * plane + src + hd);
    }
    if (want_cs) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        rx[which][0] = *reinterpret_cast<const Raw*>(planes + which * plane + src);
        rx[which][1] = *reinterpret_cast<const Raw*>(planes + which * plane + src + hd);
      }
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float g1[VEC], g2[VEC], o1[VEC], o2[VEC];
      VecIO<T, VEC>::cvt(rg[which][0], g1);
      VecIO<T, VEC>::cvt(rg[which][1], g2);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        o1[e] = g1[e] * c[e] + g2[e] * s[e];
        o2[e] = g2[e] * c[e] - g1[e] * s[e];
      }
      VecIO<T, VEC>::st(dst + which * E, o1);
      VecIO<T, VEC>::st(dst + which * E + hd, o2);
      if (want_cs) {
        float r1[VEC], r2[VEC];
        VecIO<T, VEC>::cvt(rx[which][0], r1);
        VecIO<T, VEC>::cvt(rx[which][1], r2);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float x1 = r1[e] * c[e] + r2[e] * s[e], x2 = r2[e] * c[e] - r1[e] * s[e];  // un-rotated
          acc_c[e] += g1[e] * x1 + g2[e] * x2;
          acc_s[e] += g2[e] * x1 - g1[e] * x2;
        }
      }
    }
    *reinterpret_cast<Raw*>(dst + 2 * E) = rg[2][0];
    *reinterpret_cast<Raw*>(dst + 2 * E + hd) = rg[2][1];
  }
  if (want_cs) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      atomicAdd(d_cos + tab + e, acc_c[e]);
      atomicAdd(d_sin + tab + e, acc_s[e]);
    }
  }
}

int qkv_rope_bwd(const void* d_planes, const void* planes, const float* cos_tab, const float* sin_tab,
                 void* d_qkv, float* d_cos, float* d_sin, int B, int N, int E, int H, int rope_mode,
                 int dtype, cudaStream_t st) {
  const int Dh = E / H, hd = Dh / 2;
  if (rope_mode == VRR_ROPE_NONE) d_cos = d_sin = nullptr;
  if (d_cos) {
    size_t n = (size_t)(rope_mode == VRR_ROPE_MIXED ? H : 1) * (N - 1) * hd * sizeof(float);
    if (reinterpret_cast<char*>(d_sin) == reinterpret_cast<char*>(d_cos) + n) {  // halves of one buffer: one memset node
      VRR_CUDA(cudaMemsetAsync(d_cos, 0, 2 * n, st));
    } else {
      VRR_CUDA(cudaMemsetAsync(d_cos, 0, n, st));
      VRR_CUDA(cudaMemsetAsync(d_sin, 0, n, st));
    }
  }
  const int vec = dtype == VRR_F32 ? 4 : 8;
  VRR_REQUIRE(hd % vec == 0, VRR_ERR_UNSUPPORTED, "qkv_rope_bwd: head dim %d too small for %d-wide access", Dh, vec);
  VRR_REQUIRE(((uintptr_t)d_planes & 15) == 0 && ((uintptr_t)planes & 15) == 0 && ((uintptr_t)d_qkv & 15) == 0,
              VRR_ERR_INVALID_ARG, "qkv_rope_bwd: pointers must be 16-byte aligned");
  const int threads = 128, blocks = ceil_div(H * N * (hd / vec), threads);
  int chunks = 1;
  while (chunks < B && blocks * chunks < 8 * sm_count()) chunks *= 2;
  const int b_chunk = ceil_div(B, chunks);
  dim3 grid(blocks, ceil_div(B, b_chunk));
  if (dtype == VRR_F32)
    qkv_rope_bwd_kernel<float, 4><<<grid, threads, 0, st>>>((const float*)d_planes, (const float*)planes, cos_tab,
                                                           sin_tab, (float*)d_qkv, d_cos, d_sin, B, N, E, H, Dh,
                                                           rope_mode, b_chunk);
  else
    qkv_rope_bwd_kernel<__nv_bfloat16, 8><<<grid, threads, 0, st>>>(
        (const __nv_bfloat16*)d_planes, (const __nv_bfloat16*)planes, cos_tab, sin_tab,
        (__nv_bfloat16*)d_qkv, d_cos, d_sin, B, N, E, H, Dh, rope_mode, b_chunk);
  VRR_LAUNCHED();
  return VRR_OK;
}

// cos / sin tables [heads][rows][hd] -> packed[heads][hd / 2][rows] of float4 {cos[2q], sin[2q], cos[2q+1], sin[2q+1]}:
// the layout the QKV GEMM epilogue reads.  There a LANE is a token row, so with the row-major tables one 16-byte load
// instruction touched 32 different 128-byte lines (the table row of every lane) and the L1 tag stage, not the math,
// set the epilogue's pace (ViT-B: 189 us with rope-mixed against 131 us without rotation); with rows innermost a
// load instruction covers 512 contiguous bytes.
__global__ void rope_pack_tables_kernel(const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                        float4* __restrict__ packed, int heads, int rows, int hd) {
  const int quads = hd >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * quads * rows) return;
  const int t = idx % rows, q = (idx / rows) % quads, h = idx / (rows * quads);
  const size_t src = ((size_t)h * rows + t) * hd + 2 * q;
  packed[idx] = make_float4(cos_tab[src], sin_tab[src], cos_tab[src + 1], sin_tab[src + 1]);
}

int rope_pack_tables(const float* cos_tab, const float* sin_tab, float* packed, int heads, int rows, int hd, cudaStream_t st) {
  VRR_REQUIRE(hd % 2 == 0 && ((uintptr_t)packed & 15) == 0, VRR_ERR_INVALID_ARG,
              "rope_pack_tables: half head dim must be even and `packed` 16-byte aligned");
  const int total = heads * (hd / 2) * rows;
  rope_pack_tables_kernel<<<ceil_div(total, 256), 256, 0, st>>>(cos_tab, sin_tab, reinterpret_cast<float4*>(packed), heads, rows, hd);
  VRR_LAUNCHED();
  return VRR_OK;
}

template <typename T>
__global__ void rope_apply_kernel(const T* __restrict__ q_in, const T* __restrict__ k_in,
                                  const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                  T* __restrict__ q_out, T* __restrict__ k_out, int B, int H, int Nr,
                                  int Dh, int rope_mode, int inverse) {
  const int hd = Dh >> 1;
  const size_t total = (size_t)B * H * Nr * hd;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int dd = idx % hd;
  const size_t row = idx / hd;  // (b*H + h)*Nr + t
  const int t = row % Nr, h = (row / Nr) % H;
  const size_t tab = ((size_t)(rope_mode == VRR_ROPE_MIXED ? h * Nr : 0) + t) * hd + dd;
  const float c = cos_tab[tab], s = inverse ? -sin_tab[tab] : sin_tab[tab];
  const size_t o = row * Dh + dd;
  const float q1 = Elem<T>::ld(q_in + o), q2 = Elem<T>::ld(q_in + o + hd);
  const float k1 = Elem<T>::ld(k_in + o), k2 = Elem<T>::ld(k_in + o + hd);
  Elem<T>::st(q_out + o, q1 * c - q2 * s);
  Elem<T>::st(q_out + o + hd, q1 * s + q2 * c);
  Elem<T>::st(k_out + o, k1 * c - k2 * s);
  Elem<T>::st(k_out + o + hd, k1 * s + k2 * c);
}

int rope_apply(const void* q_in, const void* k_in, const float* cos_tab, const float* sin_tab,
               void* q_out, void* k_out, int B, int H, int Nr, int Dh, int rope_mode, int inverse,
               int dtype, cudaStream_t st) {
  const size_t total = (size_t)B * H * Nr * (Dh / 2);
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (dtype == VRR_F32)
    rope_apply_kernel<float><<<blocks, threads, 0, st>>>((const float*)q_in, (const float*)k_in, cos_tab, sin_tab,
                                                        (float*)q_out, (float*)k_out, B, H, Nr, Dh, rope_mode, inverse);
  else
    rope_apply_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(
        (const __nv_bfloat16*)q_in, (const __nv_bfloat16*)k_in, cos_tab, sin_tab, (__nv_bfloat16*)q_out,
        (__nv_bfloat16*)k_out, B, H, Nr, Dh, rope_mode, inverse);
  VRR_LAUNCHED();
  return VRR_OK;
}

// Gradient of the stand-alone rotate-half w.r.t. its tables: d_cos = sum (g1 x1 + g2 x2),
// d_sin = sum (g2 x1 - g1 x2) over q and k, over the batch (and over heads for the head-shared axial
// table).  One thread per table entry, a plain loop over what it sums: deterministic, no atomics.
template <typename T>
__global__ void rope_table_grad_kernel(const T* __restrict__ q_in, const T* __restrict__ k_in,
                                       const T* __restrict__ dq, const T* __restrict__ dk, float* __restrict__ d_cos,
                                       float* __restrict__ d_sin, int B, int H, int Nr, int Dh, int rope_mode) {
  const int hd = Dh >> 1;
  const int heads_tab = rope_mode == VRR_ROPE_MIXED ? H : 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads_tab * Nr * hd) return;
  const int dd = idx % hd, t = (idx / hd) % Nr, ht = idx / (hd * Nr);
  const int h0 = rope_mode == VRR_ROPE_MIXED ? ht : 0, h1 = rope_mode == VRR_ROPE_MIXED ? ht + 1 : H;
  float ac = 0.f, as = 0.f;
  for (int b = 0; b < B; ++b)
    for (int h = h0; h < h1; ++h) {
      const size_t o = (((size_t)b * H + h) * Nr + t) * Dh + dd;
      const float q1 = Elem<T>::ld(q_in + o), q2 = Elem<T>::ld(q_in + o + hd);
      const float k1 = Elem<T>::ld(k_in + o), k2 = Elem<T>::ld(k_in + o + hd);
      const float gq1 = Elem<T>::ld(dq + o), gq2 = Elem<T>::ld(dq + o + hd);
      const float gk1 = Elem<T>::ld(dk + o), gk2 = Elem<T>::ld(dk + o + hd);
      ac += gq1 * q1 + gq2 * q2 + gk1 * k1 + gk2 * k2;
      as += gq2 * q1 - gq1 * q2 + gk2 * k1 - gk1 * k2;
    }
  d_cos[idx] = ac;
  d_sin[idx] = as;
}

int rope_table_grad(const void* q_in, const void* k_in, const void* dq, const void* dk, float* d_cos, float* d_sin,
                    int B, int H, int Nr, int Dh, int rope_mode, int dtype, cudaStream_t st) {
  const int total = (rope_mode == VRR_ROPE_MIXED ? H : 1) * Nr * (Dh / 2);
  const int threads = 128, blocks = ceil_div(total, threads);
  if (dtype == VRR_F32)
    rope_table_grad_kernel<float><<<blocks, threads, 0, st>>>((const float*)q_in, (const float*)k_in, (const float*)dq,
                                                             (const float*)dk, d_cos, d_sin, B, H, Nr, Dh, rope_mode);
  else
    rope_table_grad_kernel<__nv_bfloat16><<<blocks, threads, 0, st>>>(
        (const __nv_bfloat16*)q_in, (const __nv_bfloat16*)k_in, (const __nv_bfloat16*)dq, (const __nv_bfloat16*)dk,
        d_cos, d_sin, B, H, Nr, Dh, rope_mode);
  VRR_LAUNCHED();
  return VRR_OK;
}

}  // namespace vrr
