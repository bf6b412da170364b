// Measured FP64 FMA throughput of the device (the FP64 roofline denominator
// SURVEY.md section 8(d) asks for; MEASURED_PEAKS.json has none).  Every thread
// runs 8 independent DFMA chains; grid = 8 CTAs of 256 threads per SM.  Prints
// one JSON line.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-9, x2 = x0 + 2e-9, x3 = x0 + 3e-9;
  double x4 = x0 + 4e-9, x5 = x0 + 5e-9, x6 = x0 + 6e-9, x7 = x0 + 7e-9;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) return 1;
  const int grid = p.multiProcessorCount * 8, block = 256, iters = 4096;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * block);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_dfma<<<grid, block>>>(out, 64, 0.999999, 1e-9);   // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_dfma<<<grid, block>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fmas = (double)grid * block * iters * 16.0 * 8.0;
  const double tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"tflops\": %.3f, \"ms\": %.4f, \"sms\": %d, \"dfma_per_clk_per_sm\": %.2f, "
         "\"clock_khz_nominal\": %d, \"kernel\": \"8 independent DFMA chains per thread, "
         "%d x 256 threads\"}\n",
         tflops, best, p.multiProcessorCount,
         fmas / (best * 1e-3) / ((double)clk * 1e3) / p.multiProcessorCount, clk, grid);
  cudaFree(out);
  return cudaGetLastError() != cudaSuccess;
}
