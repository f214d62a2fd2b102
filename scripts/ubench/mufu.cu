// MUFU.EX2 throughput per SM on this GPU (the roof of the softmax exponentials): N warps per SM issuing
// independent ex2.approx chains.   nvcc -arch=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d;
  cudaMalloc(&d, 148 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int threads : {128, 256, 512, 1024}) {
    const int iters = 20000;
    k<<<148, threads>>>(d, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<148, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * threads * 8.0 * iters;
    printf("threads/SM %4d: %.1f Gex2/s total, %.2f ex2 per ns per SM (x clock GHz -> per clk: at %.2f GHz nominal = %.2f/clk/SM)\n", threads,
           ops / ms / 1e6, ops / ms / 1e6 / 148, clk / 1e6, ops / ms / 1e6 / 148 / (clk / 1e6));
  }
  return 0;
}
