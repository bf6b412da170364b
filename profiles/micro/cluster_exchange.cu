// Micro-benchmark: hierarchical all-reduce of an NT-vector over G CTAs:
// thread-block clusters of CL CTAs reduce through distributed shared memory,
// the G/CL clusters exchange through tagged slots in global memory (two hops,
// 1/CL of the volume), DSMEM all-gather.  Compare with exchange_lat.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_exchange cluster_exchange.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
struct __align__(16) Slot { uint32_t lo, t0, hi, t1; };
__device__ __forceinline__ void sstore(Slot* p, double v, uint32_t tag) {
  uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ bool sload(const Slot* p, uint32_t tag, double& v) {
  uint32_t lo, t0, hi, t1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(p));
  v = __hiloint2double((int)hi, (int)lo);
  return t0 == tag && t1 == tag;
}

template <int CL>
__global__ void k_cluster(Slot* part, Slot* box, int NT, int iters, long long* out, double* check) {
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, BT = blockDim.x;
  const int G = gridDim.x, c = blockIdx.x / CL, r = cluster.block_rank(), NC = G / CL;
  __shared__ double vals[1024];     // this CTA's d[n]
  __shared__ double cpart[256];     // cluster partial of range r
  __shared__ double eps[1024];      // gathered result
  __shared__ double mine[256];
  const int RL = 1024 / CL;         // range length per rank (NT <= CL * RL)
  const int SL = RL / NC;           // sub-slice per owner cluster
  long long t0 = clock64(), tA = 0, tB = 0, tC = 0;
  for (int it = 1; it <= iters; ++it) {
    long long a0 = clock64();
    for (int n = tid; n < 1024; n += BT) vals[n] = (n < NT) ? 1.0 + n + blockIdx.x : 0.0;
    cluster.sync();
    // reduce-scatter inside the cluster: rank r sums range r over the CL CTAs
    if (tid < RL) {
      double s = 0.0;
      for (int q = 0; q < CL; ++q) s += cluster.map_shared_rank(vals, q)[r * RL + tid];
      cpart[tid] = s;
    }
    __syncthreads();
    long long a1 = clock64();
    // global: cluster partial of (owner cluster o, ni) -> owner CTA (o, r)
    if (tid < RL) {
      const int o = tid / SL, ni = tid % SL;
      sstore(&part[(((size_t)(o * CL + r) * NC) + c) * SL + ni], cpart[tid], it);
    }
    // owner: reduce over the NC clusters
    if (tid < NC * SL) {
      const int cc = tid / SL, ni = tid % SL;
      double v;
      while (!sload(&part[(((size_t)(c * CL + r) * NC) + cc) * SL + ni], it, v)) {}
      mine[tid] = v;
    }
    __syncthreads();
    if (tid < SL) {
      double e = 0.0;
      for (int cc = 0; cc < NC; ++cc) e += mine[cc * SL + tid];
      mine[tid] = e;
    }
    __syncthreads();
    if (tid < NC * SL) {   // push to the rank-r CTA of every cluster
      const int cc = tid / SL, ni = tid % SL;
      sstore(&box[((size_t)(cc * CL + r) * NC + c) * SL + ni], mine[ni], it);
    }
    long long a2 = clock64();
    // fetch range r
    if (tid < RL) {
      double v;
      while (!sload(&box[(size_t)(c * CL + r) * RL + tid], it, v)) {}
      // all-gather through DSMEM
      for (int q = 0; q < CL; ++q) cluster.map_shared_rank(eps, q)[r * RL + tid] = v;
    }
    cluster.sync();
    long long a3 = clock64();
    tA += a1 - a0;
    tB += a2 - a1;
    tC += a3 - a2;
  }
  if (tid == 0 && blockIdx.x == 0) {
    out[0] = clock64() - t0;
    out[1] = tA;
    out[2] = tB;
    out[3] = tC;
    check[0] = eps[5];
    check[1] = eps[900];
  }
}

template <int CL>
void run(Slot* part, Slot* box, long long* out, double* check, int G) {
  const int NT = 999, iters = 200;
  cudaMemset(part, 0, (size_t)148 * 4096 * sizeof(Slot));
  cudaMemset(box, 0, (size_t)148 * 4096 * sizeof(Slot));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(256);
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  if (CL > 8) cudaFuncSetAttribute(k_cluster<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int ncl = 0;
  cudaOccupancyMaxActiveClusters(&ncl, k_cluster<CL>, &cfg);
  int nt = NT, itv = iters;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_cluster<CL>, part, box, nt, itv, out, check);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long h[4] = {0, 0, 0, 0};
  double ck[2] = {0, 0};
  cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
  cudaMemcpy(ck, check, 16, cudaMemcpyDeviceToHost);
  // expected eps[n] = sum over G CTAs of (1 + n + b) = G (1 + n) + G (G - 1) / 2
  printf("cluster %2d x %3d CTAs (max active clusters %d): %s / %s  cycles/round %lld (dsmem reduce %lld, global 2 hops %lld, fetch+gather %lld)  check %.1f (want %.1f) %.1f (want %.1f)\n",
         CL, G, ncl, cudaGetErrorString(e), cudaGetErrorString(e2), h[0] / iters, h[1] / iters, h[2] / iters,
         h[3] / iters, ck[0], G * 6.0 + G * (G - 1) / 2.0, ck[1], G * 901.0 + G * (G - 1) / 2.0);
}

int main() {
  Slot *part, *box;
  long long* out;
  double* check;
  cudaMalloc(&part, (size_t)148 * 4096 * sizeof(Slot));
  cudaMalloc(&box, (size_t)148 * 4096 * sizeof(Slot));
  cudaMalloc(&out, 64);
  cudaMalloc(&check, 64);
  run<8>(part, box, out, check, 128);
  run<8>(part, box, out, check, 128);
  run<4>(part, box, out, check, 128);
  run<16>(part, box, out, check, 128);
  return 0;
}
