// Latency of the building blocks of a sequential chain inside ONE warp on sm_100a
// (cycles per dependent link, clock64 around a loop of LINKS links).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chain_lat chain_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define LINKS 2000

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(double* out, long long* cyc, double seed, int nthreads_bar) {
  __shared__ __align__(16) double buf[2][64];
  __shared__ __align__(16) double big[32 * 2 * 12];
  __shared__ uint64_t bar;
  const int lane = threadIdx.x & 31;
  double v = seed + lane * 1e-3, w = seed * 0.5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  }
  for (int i = threadIdx.x; i < 32 * 2 * 12; i += blockDim.x) big[i] = 1e-9 * i;
  buf[0][threadIdx.x & 63] = 0.0;
  buf[1][threadIdx.x & 63] = 0.0;
  __syncthreads();
  if (MODE == 6 && threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)));
  __syncthreads();
  long long t0 = clock64();
  int cur = 0;
  for (int i = 0; i < LINKS; ++i) {
    if (MODE == 0) {   // DFMA chain
      v = fma(v, 1.0000001, w);
    } else if (MODE == 1) {   // double SHFL.BFLY + DADD
      v += __shfl_xor_sync(0xffffffffu, v, 1);
    } else if (MODE == 2) {   // STS.64 -> __syncwarp -> LDS.64 (neighbour) + DADD
      buf[cur][lane] = v;
      __syncwarp();
      v += buf[cur][lane ^ 1];
      cur ^= 1;
    } else if (MODE == 3) {   // STS.64 -> bar.sync (all threads) -> LDS.64 + DADD
      buf[cur][threadIdx.x & 63] = v;
      asm volatile("bar.sync 1, %0;" ::"r"(nthreads_bar) : "memory");
      v += buf[cur][(threadIdx.x & 63) ^ 1];
      cur ^= 1;
    } else if (MODE == 4) {   // double SHFL.IDX + DADD
      v += __shfl_sync(0xffffffffu, v, (lane + 5) & 31);
    } else if (MODE == 5) {   // 11 independent LDS.128 (prefetch batch) then shuffle link
      const double2* p = reinterpret_cast<const double2*>(big) + lane * 12;
      double2 a[11];
#pragma unroll
      for (int j = 0; j < 11; ++j) a[j] = p[j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
#pragma unroll
      for (int j = 0; j < 11; ++j) w += a[j].x * 1e-30;
    } else if (MODE == 6) {   // mbarrier.test_wait result consumed immediately
      uint32_t ok;
      asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      v += (double)ok;
      v = fma(v, 1.0000001, w);
    } else if (MODE == 7) {   // STS.128 -> __syncwarp -> 2 x LDS.128 (what a complex exchange needs)
      reinterpret_cast<double2*>(buf[cur])[lane] = make_double2(v, w);
      __syncwarp();
      const double2 x = reinterpret_cast<double2*>(buf[cur])[lane ^ 1];
      const double2 y = reinterpret_cast<double2*>(buf[cur])[lane ^ 2];
      v += x.x + y.y;
      cur ^= 1;
    } else if (MODE == 8) {   // 8 dependent DFMA (Horner J = 8) + 2 butterfly levels
#pragma unroll
      for (int j = 0; j < 8; ++j) v = fma(v, 1.0000001, w);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v + w;
}

template <int MODE>
void run(const char* name, int threads, int nbar) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * 8);
  cudaMalloc(&cyc, 8);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<1, threads>>>(out, cyc, 1.25, nbar);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-64s %8.1f cycles/link  (%s)\n", name, (double)h / LINKS, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("DFMA dependent", 32, 32);
  run<1>("double SHFL.BFLY + DADD", 32, 32);
  run<4>("double SHFL.IDX + DADD", 32, 32);
  run<2>("STS.64 + __syncwarp + LDS.64 + DADD", 32, 32);
  run<7>("STS.128 + __syncwarp + 2 LDS.128 + DADD", 32, 32);
  run<3>("STS.64 + bar.sync(32) + LDS.64 + DADD", 32, 32);
  run<3>("STS.64 + bar.sync(64: 2 warps) + LDS.64 + DADD", 64, 64);
  run<3>("STS.64 + bar.sync(160: 5 warps) + LDS.64 + DADD", 160, 160);
  run<5>("11 x LDS.128 batch + double SHFL.BFLY + DADD", 32, 32);
  run<6>("mbarrier.test_wait consumed at once + DFMA", 32, 32);
  run<8>("8 DFMA + 2 butterfly levels", 32, 32);
  return 0;
}
