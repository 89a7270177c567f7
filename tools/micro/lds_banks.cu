// Micro-benchmark: shared-memory wavefronts of the window kernel's lane -> address patterns (development aid).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_banks lds_banks.cu ; run under
// ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int pitch, int base_words) {
  extern __shared__ unsigned s[];
  for (int i = threadIdx.x; i < 57000; i += blockDim.x) s[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, lx = lane & 7, ly = lane >> 3;
  unsigned acc = 0;
  for (int it = 0; it < 256; it++) {
    int w;
    if (MODE == 0) w = base_words + it * 3 + lx + pitch * ly;                       // converged warp
    if (MODE == 1) w = base_words + it * 3 + lx + pitch * ((ly + it) % 38);          // converged, ring rows
    if (MODE == 2) w = base_words + ((lane < 16) ? it * 3 : 1520 * 5 + it * 7 + 11) + lx + pitch * ly;   // two groups
    if (MODE == 3) w = (lane * 97 + it * 131) % 57000;                              // random
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(s) + w * 4));
    acc += v;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  unsigned* d; cudaMalloc(&d, 1 << 20);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231040);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231040);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231040);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231040);
  k<0><<<1, 32, 231040>>>(d, 40, 0);     // launch 0
  k<0><<<1, 32, 231040>>>(d, 40, 1520);  // 1
  k<0><<<1, 32, 231040>>>(d, 32, 0);     // 2: pitch 32 -> 4-way expected
  k<1><<<1, 32, 231040>>>(d, 40, 0);     // 3
  k<2><<<1, 32, 231040>>>(d, 40, 0);     // 4
  k<3><<<1, 32, 231040>>>(d, 40, 0);     // 5
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
