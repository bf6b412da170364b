// Micro-benchmark: latency of the cross-CTA exchange pattern of kq_picard.cuh
// (partials -> owners reduce -> push to mailboxes -> fetch) without any compute.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_lat exchange_lat.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
struct __align__(16) Slot { uint32_t lo, t0, hi, t1; };

template <int MODE>
__device__ __forceinline__ void sstore(Slot* p, double v, uint32_t tag) {
  uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  if (MODE == 0)
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
  else
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
template <int MODE>
__device__ __forceinline__ bool sload(const Slot* p, uint32_t tag, double& v) {
  uint32_t lo, t0, hi, t1;
  if (MODE == 0)
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(p));
  else
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(p));
  v = __hiloint2double((int)hi, (int)lo);
  return t0 == tag && t1 == tag;
}

// pattern 0: full exchange (A: partials n-major, B: owners reduce + push, C: fetch mailbox)
// pattern 1: ping-pong between CTA 0 and CTA gridDim.x-1 (one slot)
template <int MODE>
__global__ void k_exchange(Slot* part, Slot* box, int NT, int stride, int iters, int pattern, int sleep_ns,
                           long long* out) {
  const int tid = threadIdx.x, BT = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = BT >> 5;
  const int G = gridDim.x, c = blockIdx.x;
  long long t0 = clock64();
  long long tB = 0, tC = 0;
  if (pattern == 2) {
  } else if (pattern == 1) {
    if (tid == 0 && (c == 0 || c == G - 1)) {
      for (int it = 1; it <= iters; ++it) {
        double v;
        if (c == 0) {
          sstore<MODE>(&box[0], 1.0, it);
          while (!sload<MODE>(&box[1], it, v)) {}
        } else {
          while (!sload<MODE>(&box[0], it, v)) {}
          sstore<MODE>(&box[1], 1.0, it);
        }
      }
    }
  } else {
    const int Wc = (NT + G - 1) / G, n_lo = c * Wc;
    for (int it = 1; it <= iters; ++it) {
      long long a0 = clock64();
      // A: partial for every n (thread owns 4 consecutive n)
      for (int w = 0; w < 4; ++w) {
        const int n = tid * 4 + w;
        if (n < NT) sstore<MODE>(&part[(size_t)n * G + c], 1.0 + n, it);
      }
      // B: owners
      for (int ni = warp; ni < Wc; ni += nw) {
        const int n = n_lo + ni;
        if (n < NT) {
          double acc = 0.0;
          for (int cb = lane; cb < G; cb += 32) {
            double v;
            while (!sload<MODE>(&part[(size_t)n * G + cb], it, v)) { if (sleep_ns) __nanosleep(sleep_ns); }
            acc += v;
          }
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          for (int cb = lane; cb < G; cb += 32) sstore<MODE>(&box[(size_t)cb * stride + n], acc, it);
        }
      }
      long long a1 = clock64();
      // C: fetch own mailbox
      double s = 0.0;
      for (int n = tid; n < NT; n += BT) {
        double v;
        while (!sload<MODE>(&box[(size_t)c * stride + n], it, v)) { if (sleep_ns) __nanosleep(sleep_ns); }
        s += v;
      }
      if (s == -1.0) out[3] = 1;
      __syncthreads();
      long long a2 = clock64();
      tB += a1 - a0;
      tC += a2 - a1;
    }
  }
  if (pattern == 2) {
    // line-coalesced variant: part[o][c][ni], box[c][o][ni], WcP = 8 slots = one 128-byte line
    __shared__ double vals[1024 + 64];
    __shared__ double red[8][8];
    __shared__ double own[8];
    const int Wc = (NT + G - 1) / G, WcP = 8, o = c;
    for (int it = 1; it <= iters; ++it) {
      long long a0 = clock64();
      // pass B stand-in: d values into smem, then full-line stores
      for (int w = 0; w < 4; ++w) { const int n = tid * 4 + w; if (n < NT) vals[n] = 1.0 + n; }
      __syncthreads();
      for (int j = tid; j < G * WcP; j += BT) {
        const int oo = j >> 3, ni = j & 7, n = oo * Wc + ni;
        if (ni < Wc && n < NT) sstore<MODE>(&part[((size_t)oo * G + c) * WcP + ni], vals[n], it);
      }
      // owners: reduce over c'
      double v4 = 0.0;
      for (int j = tid; j < G * WcP; j += BT) {
        const int ni = j & 7, n = o * Wc + ni;
        if (ni < Wc && n < NT) {
          double v;
          while (!sload<MODE>(&part[(size_t)o * G * WcP + j], it, v)) { if (sleep_ns) __nanosleep(sleep_ns); }
          v4 += v;
        }
      }
      v4 += __shfl_xor_sync(0xffffffffu, v4, 8);
      v4 += __shfl_xor_sync(0xffffffffu, v4, 16);
      if (lane < 8) red[warp][lane] = v4;
      __syncthreads();
      if (tid < 8) { double e = 0.0; for (int w = 0; w < nw; ++w) e += red[w][tid]; own[tid] = e; }
      __syncthreads();
      for (int j = tid; j < G * WcP; j += BT) {
        const int cb = j >> 3, ni = j & 7, n = o * Wc + ni;
        if (ni < Wc && n < NT) sstore<MODE>(&box[(size_t)cb * stride + o * WcP + ni], own[ni], it);
      }
      long long a1 = clock64();
      double s = 0.0;
      for (int j = tid; j < G * WcP; j += BT) {
        const int oo = j >> 3, ni = j & 7, n = oo * Wc + ni;
        if (ni < Wc && n < NT) {
          double v;
          while (!sload<MODE>(&box[(size_t)c * stride + j], it, v)) { if (sleep_ns) __nanosleep(sleep_ns); }
          s += v;
        }
      }
      if (s == -1.0) out[3] = 1;
      __syncthreads();
      long long a2 = clock64();
      tB += a1 - a0;
      tC += a2 - a1;
    }
  }
  if (tid == 0 && c == 0) {
    out[0] = clock64() - t0;
    out[1] = tB;
    out[2] = tC;
  }
}

int main(int argc, char** argv) {
  const int NT = 999, stride = 2048, iters = 200;
  Slot *part, *box;
  long long* out;
  cudaMalloc(&part, (size_t)148 * 2048 * sizeof(Slot));
  cudaMalloc(&box, (size_t)148 * 2048 * sizeof(Slot));
  cudaMalloc(&out, 64);
  for (int mode = 0; mode < 1; ++mode)
    for (int pattern = 0; pattern < 3; ++pattern)
      for (int G : {128})
        for (int sleep_ns : {0, 100}) {
          if (pattern == 1 && sleep_ns) continue;
          cudaMemset(part, 0, (size_t)148 * 2048 * sizeof(Slot));
          cudaMemset(box, 0, (size_t)148 * 2048 * sizeof(Slot));
          int nt = NT, st = stride, itv = iters, pat = pattern, sl = sleep_ns;
          void* args[] = {&part, &box, &nt, &st, &itv, &pat, &sl, &out};
          cudaError_t e = cudaLaunchCooperativeKernel(
              mode == 0 ? (const void*)k_exchange<0> : (const void*)k_exchange<1>, dim3(G), dim3(256), args, 0, 0);
          cudaDeviceSynchronize();
          long long h[4] = {0, 0, 0, 0};
          cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
          printf("mode %s pattern %s G %3d sleep %3d: %s  cycles/iter total %lld  (A+B %lld, C %lld)\n",
                 mode ? "relaxed.gpu" : "volatile   ", pattern == 2 ? "coalesced" : (pattern ? "pingpong" : "exchange"), G, sleep_ns,
                 cudaGetErrorString(e), h[0] / iters, h[1] / iters, h[2] / iters);
        }
  return 0;
}
