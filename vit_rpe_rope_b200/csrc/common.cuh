// Shared host/device helpers for libvrr_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/vrr.h"

namespace vrr {

// ---- error reporting -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_impl;

#define VRR_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::vrr::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

#define VRR_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::vrr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                       __FILE__, __LINE__);                                         \
      return VRR_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define VRR_LAUNCHED()                                   \
  do {                                                   \
    ::vrr::g_launches.fetch_add(1, std::memory_order_relaxed); \
    VRR_CUDA(cudaGetLastError());                        \
  } while (0)

int require_device();  // VRR_OK or VRR_ERR_NO_DEVICE (state cached per device ordinal)
int sm_count();        // SM count of the current device
extern std::atomic<uint64_t> g_family_launches[3];  // indexed by vrr_impl: dispatches per kernel family
bool first_use_on_device(std::atomic<unsigned char>* flags);  // flags: static array [64], one per device

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device).
#define VRR_SMEM_ATTR_ONCE(kernel, bytes)                                                              \
  do {                                                                                                 \
    static std::atomic<unsigned char> flags__[64];                                                     \
    if (::vrr::first_use_on_device(flags__))                                                           \
      VRR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
  } while (0)
#define VRR_COUNT_FAMILY(f) ::vrr::g_family_launches[(f)].fetch_add(1, std::memory_order_relaxed)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- element type traits -----------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// Load 4 consecutive elements (16B-aligned for float, 8B for bf16) as floats.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

// ---- logit bias evaluated on the fly -------------------------------------------------------------
// Device-side view of vrr_bias_desc for one head.  `lut` lives in shared memory:
//   TABLE: the head's table row, lut[i - j + N - 1]
//   POLY : lut[d] = sum_k coef[k] * d^k for d in [0, 2g-2] (power-sum order of the reference,
//          models/positional_encoding.py:145-148), 0 on the cls row/column.
struct BiasView {
  int mode, n, grid;
  const float* lut;
  __device__ __forceinline__ int index(int i, int j) const {
    if (mode == VRR_BIAS_TABLE) return i - j + n - 1;
    // POLY; caller guarantees i, j >= 1
    int pi = i - 1, pj = j - 1;
    int yi = pi % grid, xi = pi / grid, yj = pj % grid, xj = pj / grid;
    return abs(yi - yj) + abs(xi - xj);
  }
  __device__ __forceinline__ float at(int i, int j) const {
    if (mode == VRR_BIAS_NONE) return 0.f;
    if (mode == VRR_BIAS_POLY && (i == 0 || j == 0)) return 0.f;
    return lut[index(i, j)];
  }
};

// Number of LUT entries a CTA keeps in shared memory for one head.
static inline int bias_lut_len(const vrr_bias_desc* b, int n) {
  if (!b || b->mode == VRR_BIAS_NONE) return 0;
  if (b->mode == VRR_BIAS_TABLE) return 2 * n - 1;
  return 2 * b->grid - 1;
}

// Fill the per-head LUT (all threads of the CTA participate; caller syncs afterwards).
__device__ __forceinline__ void fill_bias_lut(float* lut, int mode, const float* param, int heads,
                                              int len, int grid, int n, int h) {
  if (mode == VRR_BIAS_TABLE) {
    const float* row = param + (size_t)h * len;
    for (int t = threadIdx.x; t < 2 * n - 1; t += blockDim.x) lut[t] = row[t];
  } else if (mode == VRR_BIAS_POLY) {
    const float* c = param + (size_t)(heads == 1 ? 0 : h) * len;
    for (int d = threadIdx.x; d < 2 * grid - 1; d += blockDim.x) {
      float x = (float)d, pw = 1.f, acc = 0.f;
      for (int k = 0; k < len; ++k) {
        acc = fmaf(pw, c[k], acc);
        pw *= x;
      }
      lut[d] = acc;
    }
  }
}

}  // namespace vrr
