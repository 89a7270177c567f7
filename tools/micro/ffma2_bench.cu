// Throughput of scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100) on one B200: 8 independent accumulator chains per thread,
// 1024 threads per CTA, 2 CTAs per SM.  Prints GFMA/s for both.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__global__ void __launch_bounds__(1024, 2) k_scalar(float* out, float x, float c, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x), "f"(c));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(1024, 2) k_packed(float* out, float x, float c, int iters) {
  unsigned long long a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = pk(threadIdx.x + i, threadIdx.x - i);
  const unsigned long long xx = pk(x, x + 1), cc = pk(c, c);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma2(xx, cc, a[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 296 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int rep = 0; rep < 2; rep++) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<296, 1024>>>(out, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("scalar FFMA : %.3f ms, %.1f GFMA/s\n", ms, 296.0 * 1024 * 16 * iters / ms / 1e6);
    cudaEventRecord(e0); k_packed<<<296, 1024>>>(out, 1.0001f, 0.5f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2: %.3f ms, %.1f GFMA/s\n", ms, 296.0 * 1024 * 16 * iters / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
