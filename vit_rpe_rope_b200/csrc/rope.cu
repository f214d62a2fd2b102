// RoPE rotate-half kernels that are not GEMM epilogues:
//   * backward of the fused QKV+RoPE epilogue: un-rotate the plane gradients into the token
//     layout [B*N][3E] the two projection-backward GEMMs consume, and reduce d_cos / d_sin over
//     the batch (autograd of models/rope_utils.py:22-35 and of the cat/split at vit.py:56-68);
//   * the stand-alone public apply_rotary_emb (models/rope_utils.py:3-37).
#include "common.cuh"

namespace vrr {

template <typename T>
__global__ void qkv_rope_bwd_kernel(const T* __restrict__ d_planes, const T* __restrict__ planes,
                                    const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
                                    T* __restrict__ d_qkv, float* d_cos, float* d_sin, int B, int N,
                                    int E, int H, int Dh, int rope_mode, int b_chunk) {
  const int hd = Dh >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * N * hd) return;
  const int dd = idx % hd, t = (idx / hd) % N, h = idx / (hd * N);
  const int b0 = blockIdx.y * b_chunk, b1 = min(B, b0 + b_chunk);
  const size_t plane = (size_t)B * H * N * Dh;
  const bool rot = rope_mode != VRR_ROPE_NONE && t >= 1;
  float c = 1.f, s = 0.f;
  size_t tab = 0;
  if (rot) {
    tab = ((size_t)(rope_mode == VRR_ROPE_MIXED ? h * (N - 1) : 0) + (t - 1)) * hd + dd;
    c = cos_tab[tab];
    s = sin_tab[tab];
  }
  float acc_c = 0.f, acc_s = 0.f;
  for (int b = b0; b < b1; ++b) {
    const size_t src = (((size_t)b * H + h) * N + t) * Dh + dd;
    T* dst = d_qkv + ((size_t)b * N + t) * (3 * E) + h * Dh + dd;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float g1 = Elem<T>::ld(d_planes + which * plane + src);
      const float g2 = Elem<T>::ld(d_planes + which * plane + src + hd);
      Elem<T>::st(dst + which * E, g1 * c + g2 * s);
      Elem<T>::st(dst + which * E + hd, g2 * c - g1 * s);
      if (rot && d_cos) {
        const float r1 = Elem<T>::ld(planes + which * plane + src);
        const float r2 = Elem<T>::ld(planes + which * plane + src + hd);
        const float x1 = r1 * c + r2 * s, x2 = r2 * c - r1 * s;  // un-rotated activations
        acc_c += g1 * x1 + g2 * x2;
        acc_s += g2 * x1 - g1 * x2;
      }
    }
    dst[2 * E] = d_planes[2 * plane + src];
    dst[2 * E + hd] = d_planes[2 * plane + src + hd];
  }
  if (rot && d_cos) {
    atomicAdd(d_cos + tab, acc_c);
    atomicAdd(d_sin + tab, acc_s);
  }
}

int qkv_rope_bwd(const void* d_planes, const void* planes, const float* cos_tab, const float* sin_tab,
                 void* d_qkv, float* d_cos, float* d_sin, int B, int N, int E, int H, int rope_mode,
                 int dtype, cudaStream_t st) {
  const int Dh = E / H, hd = Dh / 2;
  if (rope_mode == VRR_ROPE_NONE) d_cos = d_sin = nullptr;
  if (d_cos) {
    size_t n = (size_t)(rope_mode == VRR_ROPE_MIXED ? H : 1) * (N - 1) * hd * sizeof(float);
    VRR_CUDA(cudaMemsetAsync(d_cos, 0, n, st));
    VRR_CUDA(cudaMemsetAsync(d_sin, 0, n, st));
  }
  const int threads = 128, blocks = ceil_div(H * N * hd, threads);
  int chunks = 1;
  while (chunks < B && blocks * chunks < 4 * sm_count()) chunks *= 2;
  const int b_chunk = ceil_div(B, chunks);
  dim3 grid(blocks, ceil_div(B, b_chunk));
  if (dtype == VRR_F32)
    qkv_rope_bwd_kernel<float><<<grid, threads, 0, st>>>((const float*)d_planes, (const float*)planes, cos_tab,
                                                        sin_tab, (float*)d_qkv, d_cos, d_sin, B, N, E, H, Dh,
                                                        rope_mode, b_chunk);
  else
    qkv_rope_bwd_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(
        (const __nv_bfloat16*)d_planes, (const __nv_bfloat16*)planes, cos_tab, sin_tab,
        (__nv_bfloat16*)d_qkv, d_cos, d_sin, B, N, E, H, Dh, rope_mode, b_chunk);
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

}  // namespace vrr
