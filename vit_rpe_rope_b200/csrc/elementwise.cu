// Reductions / elementwise kernels of the projection and MLP backward (row N1 of SURVEY.md 8(f)):
//   * colsum: bias gradient db[c] = sum_m dy[m][c]   (ATen's generic reduce ran at ~1.3 TB/s here)
//   * gelu_bwd_colsum: dh = dy * gelu'(h) (exact erf GELU, timm Mlp / nn.GELU of models/vit.py:118)
//     with the fc1 bias gradient sum_m dh[m][c] accumulated in the same pass.
#include "common.cuh"
#include "kernels.h"

namespace vrr {

namespace {

constexpr int kCsThreads = 256;   // 32 column-quads x 8 row lanes
constexpr int kCsRows = 256;      // rows per CTA

// 16-byte loads, eight of them in flight per thread: thread = VEC consecutive columns (8 bf16 / 4 fp32) x rows
// rl, rl + 16, ...; CTA = 16 column groups x 16 row lanes over kCsRows rows; grid (ceil(C / (16 VEC)), ceil(M / 256)).
// (The first version read 8 bytes per thread and row with a rolled loop: 38 us average over the step's bias
// gradients where the HBM floor is 24 us.)
template <typename T>
struct CsVec;
template <>
struct CsVec<float> {
  static constexpr int kVec = 4;
  using Raw = float4;
  static __device__ __forceinline__ void add(float (&acc)[4], const float4& v) {
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  }
};
template <>
struct CsVec<__nv_bfloat16> {
  static constexpr int kVec = 8;
  using Raw = uint4;
  static __device__ __forceinline__ void add(float (&acc)[8], const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
      acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(kCsThreads) colsum_vec_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int C) {
  constexpr int VEC = CsVec<T>::kVec;
  using Raw = typename CsVec<T>::Raw;
  __shared__ float part[16][16 * VEC + 1];
  const int cg = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int c = (blockIdx.x * 16 + cg) * VEC;
  const int r0 = blockIdx.y * kCsRows + rl, r1 = min(M, blockIdx.y * kCsRows + kCsRows);
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  if (c < C) {
    const T* p = x + (size_t)r0 * C + c;
    int r = r0;
    for (; r + 7 * 16 < r1; r += 8 * 16, p += (size_t)8 * 16 * C) {
      Raw v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const Raw*>(p + (size_t)k * 16 * C);
#pragma unroll
      for (int k = 0; k < 8; ++k) CsVec<T>::add(acc, v[k]);
    }
    for (; r < r1; r += 16, p += (size_t)16 * C) CsVec<T>::add(acc, *reinterpret_cast<const Raw*>(p));
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) part[rl][cg * VEC + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 16 * VEC) {
    const int col = blockIdx.x * 16 * VEC + threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += part[k][threadIdx.x];
    if (col < C) atomicAdd(out + col, s);
  }
}

// grid (ceil(C / 128), ceil(M / kCsRows)); thread = 4 consecutive columns x strided rows
template <typename T>
__global__ void __launch_bounds__(kCsThreads) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int C) {
  __shared__ float4 part[8][32];
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + cq * 4;
  const int r0 = blockIdx.y * kCsRows, r1 = min(M, r0 + kCsRows);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (int r = r0 + rl; r < r1; r += 8) {
      const float4 v = ld4(x + (size_t)r * C + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part[rl][cq] = acc;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = part[k][cq];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(out + c + 0, acc.x);
    atomicAdd(out + c + 1, acc.y);
    atomicAdd(out + c + 2, acc.z);
    atomicAdd(out + c + 3, acc.w);
  }
}

__device__ __forceinline__ float gelu_grad(float h) {
  // d/dh [ 0.5 h (1 + erf(h / sqrt 2)) ] = Phi(h) + h * phi(h),  phi(h) = exp(-h^2/2) / sqrt(2 pi).
  // erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): erf(z) = 1 - (a1 t + .. + a5 t^5) exp(-z^2),
  // t = 1 / (1 + 0.3275911 z), z = |h| / sqrt 2 - its exponential exp(-z^2) = exp(-h^2/2) is the one
  // phi(h) needs, so the whole derivative costs one MUFU.EX2, one MUFU.RCP and ~12 FMA-pipe ops
  // (libm erff alone is ~3x that, which made this kernel issue-bound instead of HBM-bound).
  const float e = __expf(-0.5f * h * h);
  const float z = fabsf(h) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.f);           // erf(|h| / sqrt 2)
  const float cdf = 0.5f * (1.f + copysignf(erf_abs, h));  // Phi(h)
  return fmaf(h * 0.3989422804014327f, e, cdf);
}

template <typename T>
__global__ void __launch_bounds__(kCsThreads) gelu_bwd_colsum_kernel(const T* __restrict__ dy, const T* __restrict__ h,
                                                                     T* __restrict__ dh, float* __restrict__ db, int M,
                                                                     int C) {
  __shared__ float4 part[8][32];
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + cq * 4;
  const int r0 = blockIdx.y * kCsRows, r1 = min(M, r0 + kCsRows);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (int r = r0 + rl; r < r1; r += 8) {
      const size_t o = (size_t)r * C + c;
      const float4 g = ld4(dy + o), hv = ld4(h + o);
      float4 d;
      d.x = g.x * gelu_grad(hv.x);
      d.y = g.y * gelu_grad(hv.y);
      d.z = g.z * gelu_grad(hv.z);
      d.w = g.w * gelu_grad(hv.w);
      st4(dh + o, d);
      // the bias gradient sums the values as stored (rounded to T), like autograd's sum over dh
      T q[4];
      Elem<T>::st(&q[0], d.x); Elem<T>::st(&q[1], d.y); Elem<T>::st(&q[2], d.z); Elem<T>::st(&q[3], d.w);
      acc.x += Elem<T>::ld(&q[0]); acc.y += Elem<T>::ld(&q[1]); acc.z += Elem<T>::ld(&q[2]); acc.w += Elem<T>::ld(&q[3]);
    }
  }
  part[rl][cq] = acc;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 v = part[k][cq];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(db + c + 0, acc.x);
    atomicAdd(db + c + 1, acc.y);
    atomicAdd(db + c + 2, acc.z);
    atomicAdd(db + c + 3, acc.w);
  }
}

// SIMT-family fallbacks of the GEMM epilogues VRR_EPI_BIAS_GELU_GRAD / VRR_EPI_MUL (fp32, odd shapes, forced SIMT)
template <typename T>
__global__ void gelu_grad_inplace_kernel(T* __restrict__ h, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    Elem<T>::st(h + i, gelu_grad(Elem<T>::ld(h + i)));
}
template <typename T>
__global__ void mul_inplace_kernel(T* __restrict__ c, const T* __restrict__ m, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    Elem<T>::st(c + i, Elem<T>::ld(c + i) * Elem<T>::ld(m + i));
}

}  // namespace

int gelu_grad_inplace(void* h, size_t n, int dtype, cudaStream_t st) {
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (dtype == VRR_F32) gelu_grad_inplace_kernel<float><<<grid, 256, 0, st>>>((float*)h, n);
  else gelu_grad_inplace_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)h, n);
  VRR_LAUNCHED();
  return VRR_OK;
}

int mul_inplace(void* c, const void* m, size_t n, int dtype, cudaStream_t st) {
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (dtype == VRR_F32) mul_inplace_kernel<float><<<grid, 256, 0, st>>>((float*)c, (const float*)m, n);
  else mul_inplace_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)c, (const __nv_bfloat16*)m, n);
  VRR_LAUNCHED();
  return VRR_OK;
}

int colsum(const void* x, float* out, int M, int C, int dtype, cudaStream_t st) {
  VRR_REQUIRE(C % 4 == 0, VRR_ERR_UNSUPPORTED, "colsum: C = %d must be a multiple of 4", C);
  VRR_CUDA(cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), st));
  const int vec = dtype == VRR_F32 ? 4 : 8;
  if (C % vec == 0 && ((uintptr_t)x & 15) == 0) {
    dim3 grid(ceil_div(C, 16 * vec), ceil_div(M, kCsRows));
    if (dtype == VRR_F32) colsum_vec_kernel<float><<<grid, kCsThreads, 0, st>>>((const float*)x, out, M, C);
    else colsum_vec_kernel<__nv_bfloat16><<<grid, kCsThreads, 0, st>>>((const __nv_bfloat16*)x, out, M, C);
    VRR_LAUNCHED();
    return VRR_OK;
  }
  dim3 grid(ceil_div(C, 128), ceil_div(M, kCsRows));
  if (dtype == VRR_F32) colsum_kernel<float><<<grid, kCsThreads, 0, st>>>((const float*)x, out, M, C);
  else colsum_kernel<__nv_bfloat16><<<grid, kCsThreads, 0, st>>>((const __nv_bfloat16*)x, out, M, C);
  VRR_LAUNCHED();
  return VRR_OK;
}

int gelu_bwd_colsum(const void* dy, const void* h, void* dh, float* db, int M, int C, int dtype, cudaStream_t st) {
  VRR_REQUIRE(C % 4 == 0, VRR_ERR_UNSUPPORTED, "gelu_bwd: C = %d must be a multiple of 4", C);
  VRR_CUDA(cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st));
  dim3 grid(ceil_div(C, 128), ceil_div(M, kCsRows));
  if (dtype == VRR_F32)
    gelu_bwd_colsum_kernel<float><<<grid, kCsThreads, 0, st>>>((const float*)dy, (const float*)h, (float*)dh, db, M, C);
  else
    gelu_bwd_colsum_kernel<__nv_bfloat16><<<grid, kCsThreads, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)h,
                                                                      (__nv_bfloat16*)dh, db, M, C);
  VRR_LAUNCHED();
  return VRR_OK;
}

}  // namespace vrr
