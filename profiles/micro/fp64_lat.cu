// Micro-benchmark: FP64 pipe latency / issue rate and shuffle latency on
// B200 (one warp).  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_shfl(double* out, long long* cyc, int iters) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x += __shfl_xor_sync(0xffffffffu, x, 1 << (u % 5));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_bar(double* out, long long* cyc, int iters) {
  __shared__ double s[64];
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    s[(it & 1) * 32 + (threadIdx.x >> 5)] = x;
    __syncthreads();
    x += s[(it & 1) * 32 + ((threadIdx.x >> 5) ^ 1)];
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * 8);
  cudaMalloc(&cyc, 8);
  long long h;
  const int iters = 1000;
#define RUN(ILP)                                                         \
  k_dfma<ILP><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3);                  \
  k_dfma<ILP><<<1, 32>>>(out, cyc, iters, 0.999, 1e-3);                  \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                        \
  printf("DFMA ILP=%d: %.2f cycles per dependent step, %.2f cycles/instr\n", ILP, \
         (double)h / (iters * 16), (double)h / (iters * 16 * ILP));
  RUN(1) RUN(2) RUN(4) RUN(8)
  k_shfl<<<1, 32>>>(out, cyc, iters);
  k_shfl<<<1, 32>>>(out, cyc, iters);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("double SHFL.BFLY + DADD: %.2f cycles per level\n", (double)h / (iters * 16));
  k_bar<<<1, 128>>>(out, cyc, iters);
  k_bar<<<1, 128>>>(out, cyc, iters);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("STS + __syncthreads + LDS + DADD (4 warps): %.2f cycles\n", (double)h / iters);
  return 0;
}
