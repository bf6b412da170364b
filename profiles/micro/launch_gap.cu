// Micro-benchmark: what does a kernel boundary cost for a grid shaped like
// k_krotov_picard (128 CTAs x 256 threads, ~200 KB dynamic shared memory, 1 CTA per SM)?
// Launches the kernel back to back on one stream and reports
//   period        average time per launch (events over the whole train),
//   span          first CTA entry -> last CTA exit inside one launch (%globaltimer),
//   entry spread  last CTA entry - first CTA entry,
//   gap           period - span  (kernel boundary: drain + launch + CTA dispatch).
// The "pdl" rows launch with programmatic stream serialization (cudaLaunchKernelEx +
// griddepcontrol.launch_dependents / griddepcontrol.wait): the next launch's CTAs may become
// resident while the previous grid drains; overhead = period - spin time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_gap launch_gap.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int TOUCH>
__global__ void __launch_bounds__(256, 1) k_body(unsigned long long* stamps, int launch, int spin_ns,
                                                 int touch_bytes, int pdl) {
  extern __shared__ double sm[];
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const unsigned long long t0 = gtime();
  if (TOUCH) {   // touch the shared memory like the real kernel's prologue does
    for (int i = threadIdx.x; i < touch_bytes / 8; i += blockDim.x) sm[i] = (double)i;
    __syncthreads();
  }
  while (gtime() - t0 < (unsigned long long)spin_ns) {
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    stamps[((size_t)launch * gridDim.x + blockIdx.x) * 2 + 0] = t0;
    stamps[((size_t)launch * gridDim.x + blockIdx.x) * 2 + 1] = gtime();
  }
}

int main(int argc, char** argv) {
  const int grid = 128, block = 256, n = 200;
  unsigned long long* d;
  cudaMalloc(&d, sizeof(unsigned long long) * 2 * grid * n);
  std::vector<unsigned long long> h(2 * (size_t)grid * n);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  printf("%-8s %-6s %-8s %10s %10s %12s %10s\n", "launch", "smemKB", "spin_us", "period_us", "span_us",
         "entry_spread", "gap_us");
  for (int coop = 0; coop < 4; ++coop)   // 0 regular, 1 cooperative, 2 pdl, 3 pdl + cooperative
    for (int smem_kb : {0, 200})
      for (int spin_us : {5, 40}) {
        int pdl = coop >= 2;
        int smem = smem_kb * 1024, spin_ns = spin_us * 1000, touch = smem;
        cudaFuncSetAttribute(k_body<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
          cudaStreamSynchronize(st);
          cudaEventRecord(e0, st);
          for (int i = 0; i < n; ++i) {
            void* params[] = {(void*)&d, (void*)&i, (void*)&spin_ns, (void*)&touch, (void*)&pdl};
            if (coop == 1)
              cudaLaunchCooperativeKernel((const void*)k_body<1>, dim3(grid), dim3(block), params, smem, st);
            else if (coop == 0)
              cudaLaunchKernel((const void*)k_body<1>, dim3(grid), dim3(block), params, smem, st);
            else {
              cudaLaunchConfig_t cfg = {};
              cfg.gridDim = dim3(grid);
              cfg.blockDim = dim3(block);
              cfg.dynamicSmemBytes = smem;
              cfg.stream = st;
              cudaLaunchAttribute at[2];
              at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
              at[0].val.programmaticStreamSerializationAllowed = 1;
              at[1].id = cudaLaunchAttributeCooperative;
              at[1].val.cooperative = 1;
              cfg.attrs = at;
              cfg.numAttrs = (coop == 3) ? 2 : 1;
              cudaLaunchKernelExC(&cfg, (const void*)k_body<1>, params);
            }
          }
          cudaEventRecord(e1, st);
          cudaStreamSynchronize(st);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        double span = 0, spread = 0;
        for (int i = n / 2; i < n; ++i) {
          unsigned long long lo = ~0ull, hi = 0, ehi = 0;
          for (int c = 0; c < grid; ++c) {
            lo = std::min(lo, h[((size_t)i * grid + c) * 2]);
            ehi = std::max(ehi, h[((size_t)i * grid + c) * 2]);
            hi = std::max(hi, h[((size_t)i * grid + c) * 2 + 1]);
          }
          span += (double)(hi - lo);
          spread += (double)(ehi - lo);
        }
        span /= (n - n / 2) * 1e3;
        spread /= (n - n / 2) * 1e3;
        const double period = ms * 1e3 / n;
        printf("%-8s %-6d %-8d %10.2f %10.2f %12.2f %10.2f\n", coop == 0 ? "regular" : coop == 1 ? "coop" : coop == 2 ? "pdl" : "pdl+coop", smem_kb, spin_us,
               period, span, spread, period - span);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
      }
  return 0;
}
