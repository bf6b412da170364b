// Shared device helpers for the Krotov sweep kernels (sm_100a).
//
// Numerics: everything is complex128 / float64 on the FP64 CUDA cores
// (tcgen05 has no FP64 kind).  The single-step propagator
// exp(f*A*dt) v of the reference's propagators.expm
// (/root/reference/src/krotov/propagators.py:79-122) is evaluated as a
// truncated Taylor series in Horner form applied to the *vector*,
//     y = v + (h/1) fA (v + (h/2) fA (v + ... (v + (h/m) fA v))),  h = dt/s,
// repeated s times, with s = ceil(||A|| dt) and m picked from a per-binade
// table so that the first omitted term is below 2^-56 (kq_taylor_tables()).
// This needs m matvecs (O(m N^2)) instead of the O(N^3) Pade matrix
// exponential and agrees with scipy.linalg.expm to ~1e-16 per step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

#define KQ_MMAX_SMALL 5   // terms kept in registers by thread-per-objective kernels
#define KQ_LMAX 8         // pulses handled per sweep
#define KQ_MAX_BLOCKS 4096  // cross-CTA exchange slots are sized for this many CTAs
#define KQ_MAX_WORLD 16
// the per-rank (cross-GPU) slots [2][world][KQ_LMAX] start at the beginning of
// each rank's IPC exchange buffer (kq_comm.slots[r])
#define KQ_RANK_SLOT_OFFSET ((size_t)0)
#define KQ_NTC 512   // time steps of per-step scalars staged per chunk (kq_spec.cuh)
#define KQ_RING 4    // depth of the TMA ring for state rows (kq_spec.cuh)
#define KQ_TAYLOR_BINS 64
#define KQ_TAYLOR_MAXM 32
#define KQ_INVFACT_N 40

struct KqTables {
  int m_of_bin[KQ_TAYLOR_BINS];     // Taylor degree for xs in (2^-(i+1), 2^-i]
  double inv[KQ_TAYLOR_MAXM + 1];   // 1/j
  double invfact[KQ_INVFACT_N];     // 1/n!  (closed-form N = 2 step, kq_picard.cuh)
  // quarter-binade resolution (kq_lanes.cuh): degree for xs <= 2^-(i/4+1) (1 + (i%4 + 1) / 4)
  unsigned char m_fine[4 * KQ_TAYLOR_BINS];
};

// Every translation unit has its own copy of the tables, uploaded once per
// device by its kq_tables_upload_* function (kq_host.cuh).
static __constant__ KqTables c_kq_tables;

__device__ __forceinline__ cplx c_make(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ cplx c_zero() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ cplx c_add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx c_sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
// acc + a*b
__device__ __forceinline__ cplx c_fma(cplx a, cplx b, cplx acc) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
  return acc;
}
// acc + s*b (s real)
__device__ __forceinline__ cplx c_fma_real(double s, cplx b, cplx acc) {
  acc.x = fma(s, b.x, acc.x);
  acc.y = fma(s, b.y, acc.y);
  return acc;
}
// Im(conj(a) * b)
__device__ __forceinline__ double c_im_conj_mul(cplx a, cplx b) {
  return fma(a.x, b.y, -a.y * b.x);
}
// conj(a)*b accumulated
__device__ __forceinline__ cplx c_fma_conj(cplx a, cplx b, cplx acc) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.y, b.x, acc.y);
  return acc;
}

// Equation-of-motion factor f applied to w (propagators.py:94-105):
//   FSEL 0: f = -i (Hilbert space, forward)   FSEL 1: f = +i (Hilbert, backward)
//   FSEL 2: f = 1  (super-operator, both directions)
template <int FSEL>
__device__ __forceinline__ cplx apply_f(cplx w) {
  if (FSEL == 0) return make_double2(w.y, -w.x);
  if (FSEL == 1) return make_double2(-w.y, w.x);
  return w;
}

__device__ __forceinline__ double warp_allreduce_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_allreduce_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Scaling count s and Taylor degree m for a step with ||A|| dt <= x.
__device__ __forceinline__ void taylor_plan(const KqTables& T, double x, int& s, int& m,
                                            double& xs) {
  s = 1;
  xs = x;
  if (!(x <= 1.0e6)) {
    // non-finite (NaN pulses, e.g. from a zero chi norm) or absurdly large: the result is
    // garbage either way (the host checks the pulses for finiteness) -- keep the work
    // bounded instead of scaling a million times per step
    xs = 1.0;
  } else if (x > 1.0) {
    const double sd = ceil(x);
    s = (int)sd;
    xs = x / sd;
  }
  int e = (__double2hiint(xs) >> 20) & 0x7ff;   // biased exponent
  int bin = 1022 - e;                           // xs <= 2^-bin
  bin = max(0, min(KQ_TAYLOR_BINS - 1, bin));
  m = T.m_of_bin[bin];
}

// The same with the degree picked per QUARTER binade (the chain kernels of kq_lanes.cuh pay
// for every Taylor term with a round of shuffles: 10-15 % fewer terms on average).
__device__ __forceinline__ void taylor_plan_fine(const KqTables& T, double x, int& s, int& m) {
  double xs;
  taylor_plan(T, x, s, m, xs);
  const int hi = __double2hiint(xs);
  const int bin = 1022 - ((hi >> 20) & 0x7ff);
  if (bin >= 0) m = T.m_fine[4 * min(KQ_TAYLOR_BINS - 1, bin) + ((hi >> 18) & 3)];
}

// ---- low-latency flag+data exchange (two 8-byte halves, each carrying the
// tag; 8-byte stores are single-copy atomic, so a reader that sees both tags
// has the whole double -- the scheme NCCL's LL protocol uses) -------------
struct __align__(16) KqSlot { uint32_t lo, t0, hi, t1; };

__device__ __forceinline__ void slot_store(KqSlot* p, double v, uint32_t tag) {
  uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag),
               "r"(hi), "r"(tag)
               : "memory");
}
__device__ __forceinline__ bool slot_try_load(const KqSlot* p, uint32_t tag, double& v) {
  uint32_t lo, t0, hi, t1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1)
               : "l"(p)
               : "memory");
  if (t0 == tag && t1 == tag) {
    v = __hiloint2double((int)hi, (int)lo);
    return true;
  }
  return false;
}
// Wait for NB slots at once: all loads of a polling round are in flight together
// (one L2 round trip per round instead of one per slot).  Inactive entries must
// still point to readable memory.
template <int NB>
__device__ __forceinline__ void slot_wait_batch(const KqSlot* const (&p)[NB],
                                                const bool (&act)[NB], uint32_t tag,
                                                double (&v)[NB], bool& failed) {
#pragma unroll
  for (int u = 0; u < NB; ++u) v[u] = 0.0;
  if (failed) return;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t lo[NB], t0[NB], hi[NB], t1[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u)
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(lo[u]), "=r"(t0[u]), "=r"(hi[u]), "=r"(t1[u])
                   : "l"(p[u]));
    bool all = true;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const bool ok = !act[u] || (t0[u] == tag && t1[u] == tag);
      all = all && ok;
      if (act[u]) v[u] = __hiloint2double((int)hi[u], (int)lo[u]);
    }
    if (all) return;
  }
  failed = true;
}
// Spin until the slot carries `tag`; gives up (sets *failed) after ~2^24 polls
// so that a lost peer cannot hang the device.
__device__ __forceinline__ double slot_wait(const KqSlot* p, uint32_t tag, bool& failed) {
  double v = 0.0;
  if (failed) return v;
  for (int spin = 0; spin < (1 << 24); ++spin) {
    if (slot_try_load(p, tag, v)) return v;
  }
  failed = true;
  return 0.0;
}
