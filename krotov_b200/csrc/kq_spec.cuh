// Specialised thread-per-objective kernels for the dominant problem shape:
// one drift term + one control term (M = 2) driven by a single pulse (L = 1),
// N <= 4 -- BASELINE configs C1..C4 all have this shape.
//
// The sweep is a chain of nt-1 dependent steps, so the design goal is the
// shortest possible per-step instruction stream on the critical path:
//  * generator terms (pre-multiplied by the equation-of-motion factor f) sit
//    in registers for N <= 3 (shared memory for N = 4);
//  * per-step scalars (dt, guess pulse, S/lambda, sigma) are staged in shared
//    memory in chunks of KQ_NTC steps, so no global-load latency and no
//    division sits on the chain;
//  * the backward states chi(t_n) (and Phi0 for second order) are streamed
//    HBM -> shared memory by TMA bulk copies (cp.async.bulk + mbarrier) into a
//    ring KQ_RING steps deep: one contiguous K_cta*N*16-byte row per step
//    thanks to the time-major [nt][K][N] layout;
//  * the Horner/Taylor recurrence is entered through a fall-through switch on
//    a warp-uniform degree with 1/j as immediates; the degree is planned from
//    the *guess* pulse before the cross-objective reduction and verified after
//    the step (a violation replays the step with the exact plan);
//  * <chi| mu |phi> is evaluated as <mu^dag chi | phi>, with mu^dag chi (times
//    ||chi||) prepared one step ahead, so only an N-term dot product sits on
//    the chain (first order).
#pragma once
#include "kq_common.cuh"
#include "kq_small.cuh"


// ---- mbarrier / TMA bulk-copy primitives (PTX) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "KQ_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra KQ_DONE_%=;\n"
      "bra KQ_WAIT_%=;\n"
      "KQ_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Generator element type G: `cplx` in general; `double` when every generator
// term is a real matrix in Hilbert space -- then f*A = -+i*A_re is purely
// imaginary, the element r stands for i*r and a product costs 2 FMAs, not 4.
__device__ __forceinline__ cplx g_fma(cplx a, cplx y, cplx acc) { return c_fma(a, y, acc); }
__device__ __forceinline__ cplx g_fma(double r, cplx y, cplx acc) {
  acc.x = fma(-r, y.y, acc.x);   // (i r)(y.x + i y.y) = -r y.y + i r y.x
  acc.y = fma(r, y.x, acc.y);
  return acc;
}
template <int FSEL>
__device__ __forceinline__ cplx g_load(cplx t, cplx) { return apply_f<FSEL>(t); }
template <int FSEL>
__device__ __forceinline__ double g_load(cplx t, double) {
  return (FSEL == 0) ? -t.x : t.x;   // -i*t = i*(-t) ; +i*t = i*t   (t real)
}
__device__ __forceinline__ cplx g_axpy(double h, cplx a, double heps, cplx b) {
  return make_double2(fma(heps, b.x, h * a.x), fma(heps, b.y, h * a.y));
}
__device__ __forceinline__ double g_axpy(double h, double a, double heps, double b) {
  return fma(heps, b, h * a);
}

// y <- v + invj * (At y)   for At = h f A (column-major, registers)
template <int N, typename G>
__device__ __forceinline__ void horner_step(const G (&At)[N * N], const cplx (&v)[N],
                                            cplx (&y)[N], double invj) {
  cplx w[N];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    cplx a0 = c_zero(), a1 = c_zero();  // two partial sums shorten the chain
#pragma unroll
    for (int c = 0; c < N; ++c) {
      if (c & 1)
        a1 = g_fma(At[c * N + r], y[c], a1);
      else
        a0 = g_fma(At[c * N + r], y[c], a0);
    }
    w[r] = (N > 1) ? c_add(a0, a1) : a0;
  }
#pragma unroll
  for (int r = 0; r < N; ++r) y[r] = c_fma_real(invj, w[r], v[r]);
}

// y <- exp(At) v with a compile-time Taylor degree M (fully unrolled, 1/j immediates).
template <int N, int M, typename G>
__device__ __forceinline__ void expmv_fixed(const G (&At)[N * N], const cplx (&v)[N],
                                            cplx (&y)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = v[i];
#pragma unroll
  for (int j = M; j >= 1; --j) horner_step<N, G>(At, v, y, 1.0 / (double)j);
}

// y <- exp(At)^s v, run-time degree (generic path: s > 1 or m > KQ_MFIX)
template <int N, typename G>
__device__ __forceinline__ void expmv_generic(const G (&At)[N * N], const cplx (&v)[N],
                                              cplx (&y)[N], int s, int m) {
  cplx in[N];
#pragma unroll
  for (int i = 0; i < N; ++i) in[i] = v[i];
  for (int rep = 0; rep < s; ++rep) {
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = in[i];
    for (int j = m; j >= 1; --j) horner_step<N, G>(At, in, y, c_kq_tables.inv[j]);
#pragma unroll
    for (int i = 0; i < N; ++i) in[i] = y[i];
  }
}

// ---- closed form for N = 2 with a real generator matrix --------------------
// exp(i R) = e^{i t} [ C(z) I + i S(z) R' ],  t = tr(R)/2, R' = R - t I,
// z = -det R' (R'^2 = z I), C(z) = sum_j (-z)^j/(2j)!, S(z) = sum_j (-z)^j/(2j+1)!.
// The series are the even/odd parts of the same Taylor series the generic path
// sums, truncated one order later (2P >= m+3, because |z| <= 2 ||R||_1^2), so the
// accuracy is at least that of the generic path;
// the cost drops from m complex matvecs to two real Horner recurrences.
__host__ __device__ constexpr double kq_inv_fact(int n) {
  double f = 1.0;
  for (int i = 2; i <= n; ++i) f *= (double)i;
  return 1.0 / f;
}
template <int P>
__device__ __forceinline__ void expmv2_real(const double (&R)[4], const cplx (&v)[2],
                                            cplx (&y)[2]) {
  const double a = R[0], c = R[1], b = R[2], d = R[3];   // column-major [[a, b], [c, d]]
  const double t = 0.5 * (a + d), dl = 0.5 * (a - d);
  const double z = fma(dl, dl, b * c);
  double C = kq_inv_fact(2 * (P - 1)), S = kq_inv_fact(2 * (P - 1) + 1);
#pragma unroll
  for (int j = P - 2; j >= 0; --j) {
    C = fma(-z, C, kq_inv_fact(2 * j));
    S = fma(-z, S, kq_inv_fact(2 * j + 1));
  }
  // w = R' v
  const cplx w0 = make_double2(fma(dl, v[0].x, b * v[1].x), fma(dl, v[0].y, b * v[1].y));
  const cplx w1 = make_double2(fma(-dl, v[1].x, c * v[0].x), fma(-dl, v[1].y, c * v[0].y));
  // u = C v + i S w
  cplx u0 = make_double2(fma(-S, w0.y, C * v[0].x), fma(S, w0.x, C * v[0].y));
  cplx u1 = make_double2(fma(-S, w1.y, C * v[1].x), fma(S, w1.x, C * v[1].y));
  if (t != 0.0) {   // traceless generators (two-level systems) skip the phase
    double st, ct;
    sincos(t, &st, &ct);
    u0 = make_double2(fma(-st, u0.y, ct * u0.x), fma(st, u0.x, ct * u0.y));
    u1 = make_double2(fma(-st, u1.y, ct * u1.x), fma(st, u1.x, ct * u1.y));
  }
  y[0] = u0;
  y[1] = u1;
}

// One propagation step with a compile-time Taylor degree M: the closed form for
// (N = 2, real generator), the unrolled Horner recurrence otherwise.
template <int N, int M, typename G>
__device__ __forceinline__ void step_fixed(const G (&At)[N * N], const cplx (&v)[N],
                                           cplx (&y)[N]) {
  expmv_fixed<N, M, G>(At, v, y);
}
template <>
__device__ __forceinline__ void step_fixed<2, 1, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<3>(At, v, y); }
template <>
__device__ __forceinline__ void step_fixed<2, 2, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<3>(At, v, y); }
template <>
__device__ __forceinline__ void step_fixed<2, 3, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<3>(At, v, y); }
template <>
__device__ __forceinline__ void step_fixed<2, 4, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<4>(At, v, y); }
template <>
__device__ __forceinline__ void step_fixed<2, 5, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<4>(At, v, y); }
template <>
__device__ __forceinline__ void step_fixed<2, 6, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<5>(At, v, y); }
template <>
__device__ __forceinline__ void step_fixed<2, 7, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<5>(At, v, y); }
template <>
__device__ __forceinline__ void step_fixed<2, 8, double>(const double (&At)[4], const cplx (&v)[2], cplx (&y)[2]) { expmv2_real<6>(At, v, y); }

#define KQ_MFIX 8   // Taylor degrees 1..KQ_MFIX have fully unrolled step bodies

// Plan for a norm bound x: (s, m) and the largest x for which it stays valid.
__device__ __forceinline__ void plan_bound(double x, int& s, int& m, double& bound) {
  double xs;
  taylor_plan(c_kq_tables, x, s, m, xs);
  const int e = (__double2hiint(xs) >> 20) & 0x7ff;
  double top = __hiloint2double((e + 1) << 20, 0);  // top of the binade of xs
  if (e >= 1022) top = 1.0;                         // xs <= 1 always
  if (e == 0) top = 0.0;
  bound = top * (double)s;
}

// CTA-wide maximum of a non-negative double (all threads call; uses `scratch[32]`).
__device__ __forceinline__ double block_max(double v, double* scratch) {
  v = warp_allreduce_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = scratch[0];
  for (int w = 1; w < nw; ++w) r = fmax(r, scratch[w]);
  __syncthreads();
  return r;
}

// Terms of one objective, pre-rotated by f: T0f = f*T0, T1f = f*T1.
template <int N, bool INREG, typename G>
struct SpecTerms {
  G t0[INREG ? N * N : 1], t1[INREG ? N * N : 1];
  const G* s0;
  const G* s1;
  int stride;
  template <int FSEL>
  __device__ __forceinline__ void load(const cplx* g0, const cplx* g1, void* smem, int BT,
                                       int tid) {
    if (INREG) {
#pragma unroll
      for (int e = 0; e < N * N; ++e) {
        t0[e] = g_load<FSEL>(g0[e], G());
        t1[e] = g_load<FSEL>(g1[e], G());
      }
    } else {
      stride = BT;
      G* p0 = reinterpret_cast<G*>(smem) + tid;
      G* p1 = reinterpret_cast<G*>(smem) + (size_t)N * N * BT + tid;
      for (int e = 0; e < N * N; ++e) {
        p0[(size_t)e * BT] = g_load<FSEL>(g0[e], G());
        p1[(size_t)e * BT] = g_load<FSEL>(g1[e], G());
      }
      s0 = p0;
      s1 = p1;
    }
  }
  // At = h*T0f + (h*eps)*T1f
  __device__ __forceinline__ void assemble(double h, double heps, G (&At)[N * N]) const {
#pragma unroll
    for (int e = 0; e < N * N; ++e) {
      const G a = INREG ? t0[e] : s0[(size_t)e * stride];
      const G b = INREG ? t1[e] : s1[(size_t)e * stride];
      At[e] = g_axpy(h, a, heps, b);
    }
  }
};

// ---------------------------------------------------------------------------
// Propagation sweeps (optimize.py:806-886), M = 2.
// shared: [sdt KQ_NTC][spulse KQ_NTC][splan KQ_NTC bytes][scratch 32][terms (N = 4)]
//
// NVEC = 1: one state per objective.  NVEC = N: the N basis vectors are
// propagated together (pass 1 of the time-parallel sweep: the result is the
// propagator of the segment).  blockIdx.y selects the segment.
template <int N, bool INREG, int NVEC, typename G>
struct PropCtx {
  SpecTerms<N, INREG, G> T;
  cplx y[NVEC][N];
  const double* sdt;
  const double* sp;
  const unsigned char* splan;
  const double* pulse;   // global pulse row when not staged
  bool store;            // states are stored
  size_t kofs;           // k*N
  size_t row_stride;     // K*N
  bool driven, staged, valid;
  double c1_fixed, opn0, opn1;
  int base, dir;         // chunk base index, +1 forward / -1 backward
  const KqSweepArgs* args;
};

// Run consecutive steps j (moving by c.dir) while their planned degree is MT.
template <int N, bool INREG, int NVEC, typename G, int MT>
__device__ __forceinline__ int prop_run(PropCtx<N, INREG, NVEC, G>& c, int j, int jend) {
  constexpr int NN = N * N;
  while (j != jend && c.splan[j] == MT) {
    const int n = c.base + j;
    const double dtn = c.sdt[j];
    const double eps = c.driven ? (c.staged ? c.sp[j] : c.pulse[n]) : c.c1_fixed;
    G At[NN];
    int s = 1, m = MT;
    if (MT > 0) {
      c.T.assemble(dtn, dtn * eps, At);
    } else {
      double bound;
      plan_bound(dtn * fma(fabs(eps), c.opn1, c.opn0), s, m, bound);
      const double h = dtn / (double)s;
      c.T.assemble(h, h * eps, At);
    }
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
      cplx out[N];
      if (MT > 0)
        step_fixed<N, (MT > 0 ? MT : 1), G>(At, c.y[v], out);
      else
        expmv_generic<N, G>(At, c.y[v], out, s, m);
#pragma unroll
      for (int i = 0; i < N; ++i) c.y[v][i] = out[i];
    }
    if (NVEC == 1 && c.store && c.valid) {
      const size_t row = (c.dir < 0) ? (size_t)n : (size_t)n + 1;
#pragma unroll
      for (int i = 0; i < N; ++i)
        kq_store(*c.args, row * c.row_stride + c.kofs + i, c.y[0][i]);
    }
    j += c.dir;
  }
  return j;
}

template <int N, int FSEL, int NVEC, typename G>
__global__ void __launch_bounds__(256, 1) k_prop_spec(const KqSweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NN = N * N;
  constexpr bool INREG = (N <= 3);
  const int BT = blockDim.x, tid = threadIdx.x;
  const int K = a.K, NT = a.NT;
  int k = a.k_lo + blockIdx.x * BT + tid;
  const bool valid = k < a.k_lo + a.k_cnt;
  if (!valid) k = a.k_lo + a.k_cnt - 1;  // shadow thread: keeps the CTA converged
  double* sdt = reinterpret_cast<double*>(smem_raw);
  double* sp = sdt + KQ_NTC;
  unsigned char* splan = reinterpret_cast<unsigned char*>(sp + KQ_NTC);
  double* scratch = reinterpret_cast<double*>(splan + KQ_NTC);
  PropCtx<N, INREG, NVEC, G> c;
  c.T.template load<FSEL>(a.ops + ((size_t)k * 2 + 0) * NN, a.ops + ((size_t)k * 2 + 1) * NN,
                          scratch + 32, BT, tid);
  c.opn0 = a.op_norm[k * 2 + 0];
  c.opn1 = a.op_norm[k * 2 + 1];
  const int l = a.term2pulse[k * 2 + 1];
  c.driven = l >= 0;
  c.c1_fixed = (l == -1) ? 1.0 : 0.0;
  c.staged = (a.L == 1);
  c.pulse = c.driven ? a.pulses + (size_t)l * NT : nullptr;
  c.valid = valid;
  c.sdt = sdt;
  c.sp = sp;
  c.splan = splan;
  c.store = a.store != nullptr;
  c.kofs = (size_t)k * N;
  c.row_stride = (size_t)K * N;
  c.args = &a;
  c.dir = a.backward ? -1 : 1;
  // CTA-wide norm bounds for the per-step Taylor plan
  const double O0 = block_max(c.opn0, scratch);
  const double O1 = block_max(c.driven ? c.opn1 : 0.0, scratch);
  const double Oc = block_max(c.driven ? 0.0 : c.c1_fixed * c.opn1, scratch);
  // time window [w0, w1) of this CTA: the whole sweep, or segment blockIdx.y
  const int seg = blockIdx.y;
  int w0 = 0, w1 = NT;
  if (a.seg_pass) {
    if (a.backward) {
      w1 = NT - seg * a.seg_len;
      w0 = max(0, w1 - a.seg_len);
    } else {
      w0 = seg * a.seg_len;
      w1 = min(NT, w0 + a.seg_len);
    }
  }
  const bool first_seg = (seg == 0), last_seg = (seg == (int)gridDim.y - 1);
  if (NVEC > 1) {
#pragma unroll
    for (int v = 0; v < NVEC; ++v)
#pragma unroll
      for (int i = 0; i < N; ++i) c.y[v][i] = c_make(v == i ? 1.0 : 0.0, 0.0);
  } else {
    const cplx* init = a.seg_pass ? a.seg_B + ((size_t)seg * K + k) * N
                                  : a.state0 + (size_t)k * N;
#pragma unroll
    for (int i = 0; i < N; ++i) c.y[0][i] = init[i];
    if (a.store && valid && first_seg) {
      const size_t row = a.backward ? (size_t)NT : 0;
#pragma unroll
      for (int i = 0; i < N; ++i) kq_store(a, row * c.row_stride + c.kofs + i, c.y[0][i]);
    }
  }
  const int c0 = w0 / KQ_NTC, c1 = (w1 - 1) / KQ_NTC;   // chunks touched by the window
  for (int cc = 0; cc <= c1 - c0; ++cc) {
    const int ch = a.backward ? c1 - cc : c0 + cc;
    const int base = max(w0, ch * KQ_NTC);
    const int len = min(w1, (ch + 1) * KQ_NTC) - base;
    __syncthreads();
    for (int i = tid; i < len; i += BT) {
      const double dti = a.dt[base + i];
      sdt[i] = dti;
      double xmax;
      if (c.staged) {
        const double e = a.pulses[base + i];
        sp[i] = e;
        xmax = dti * (fma(fabs(e), O1, O0) + Oc);
      } else {
        double emax = 0.0;
        for (int ll = 0; ll < a.L; ++ll) emax = fmax(emax, fabs(a.pulses[(size_t)ll * NT + base + i]));
        xmax = dti * (fma(emax, O1, O0) + Oc);
      }
      int s, m;
      double bound;
      plan_bound(xmax, s, m, bound);
      splan[i] = (s == 1 && m <= KQ_MFIX) ? (unsigned char)m : (unsigned char)0;
    }
    __syncthreads();
    c.base = base;
    int j = a.backward ? len - 1 : 0;
    const int jend = a.backward ? -1 : len;
    while (j != jend) {
      switch (splan[j]) {
        case 1: j = prop_run<N, INREG, NVEC, G, 1>(c, j, jend); break;
        case 2: j = prop_run<N, INREG, NVEC, G, 2>(c, j, jend); break;
        case 3: j = prop_run<N, INREG, NVEC, G, 3>(c, j, jend); break;
        case 4: j = prop_run<N, INREG, NVEC, G, 4>(c, j, jend); break;
        case 5: j = prop_run<N, INREG, NVEC, G, 5>(c, j, jend); break;
        case 6: j = prop_run<N, INREG, NVEC, G, 6>(c, j, jend); break;
        case 7: j = prop_run<N, INREG, NVEC, G, 7>(c, j, jend); break;
        case 8: j = prop_run<N, INREG, NVEC, G, 8>(c, j, jend); break;
        default: j = prop_run<N, INREG, NVEC, G, 0>(c, j, jend); break;
      }
    }
  }
  if (NVEC > 1) {
    // propagator of the segment, column-major: column v = image of e_v
    if (valid) {
      cplx* P = a.seg_P + ((size_t)seg * K + k) * NN;
#pragma unroll
      for (int v = 0; v < NVEC; ++v)
#pragma unroll
        for (int i = 0; i < N; ++i) P[v * N + i] = c.y[v][i];
    }
  } else if (a.stateT && valid && last_seg) {
#pragma unroll
    for (int i = 0; i < N; ++i) a.stateT[(size_t)k * N + i] = c.y[0][i];
  }
}

// Boundary states of the segments: B[0] = state0, B[q+1] = P_q B[q].
// The segment propagators do not depend on the chain: they are loaded in
// batches (independent loads in flight together), so a step costs one small
// matvec instead of one global-memory round trip.
template <int N>
__global__ void k_seg_chain(const KqSweepArgs a, int nseg) {
  constexpr int NN = N * N;
  constexpr int NB = (N <= 2) ? 8 : (N == 3 ? 4 : 2);   // propagators per batch (registers)
  const int k = a.k_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.k_lo + a.k_cnt) return;
  const int K = a.K;
  const cplx* __restrict__ segP = a.seg_P;
  cplx* __restrict__ segB = a.seg_B;
  cplx b[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    b[i] = a.state0[(size_t)k * N + i];
    segB[((size_t)0 * K + k) * N + i] = b[i];
  }
  for (int q0 = 0; q0 < nseg; q0 += NB) {
    cplx P[NB][NN];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int q = min(q0 + u, nseg - 1);
#pragma unroll
      for (int e = 0; e < NN; ++e) P[u][e] = segP[((size_t)q * K + k) * NN + e];
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      if (q0 + u < nseg) {
        cplx o[N];
#pragma unroll
        for (int r = 0; r < N; ++r) {
          cplx acc = c_zero();
#pragma unroll
          for (int c = 0; c < N; ++c) acc = c_fma(P[u][c * N + r], b[c], acc);
          o[r] = acc;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          b[i] = o[i];
          segB[((size_t)(q0 + u + 1) * K + k) * N + i] = o[i];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Fused update + forward sweep (optimize.py:449-500), M = 2, L = 1.
// shared: [red 2*32 | tot 2 | pad][mbar KQ_RING][sdt|sg|ssl|ssig|sbound KQ_NTC each]
//         [splan KQ_NTC bytes][chi ring KQ_RING * BT*N][Phi0 ring (SECOND)][terms (N = 4)]
template <int N, bool INREG, bool SECOND, typename G>
struct FwCtx {
  SpecTerms<N, INREG, G> T;
  cplx phi[N], chi[N], eta[N], dphi[N], mu[N * N];
  double cnorm, opn0, opn1, c1_fixed, ga, O0, O1, Oc;
  bool driven, valid, writer, multi, failed;
  int lane, warp, nwarps, nblk, tl, last_row, base, K, k;
  double *red, *tot;
  const double *sdt, *sg, *ssl, *ssig, *sbound;
  const unsigned char* splan;
  uint64_t* mbar;
  cplx *ring, *ring0;
  size_t stage_cplx;
  uint32_t row_bytes;
  int k0;
};

template <int N, bool INREG, bool SECOND, typename G>
__device__ __forceinline__ void fw_make_eta(FwCtx<N, INREG, SECOND, G>& c) {
#pragma unroll
  for (int cc = 0; cc < N; ++cc) {
    cplx acc = c_zero();
#pragma unroll
    for (int r = 0; r < N; ++r) acc = c_fma_conj(c.mu[cc * N + r], c.chi[r], acc);
    c.eta[cc] = make_double2(acc.x * c.cnorm, acc.y * c.cnorm);
  }
}

template <int N, bool INREG, bool SECOND, typename G>
__device__ __forceinline__ void fw_issue_row(const KqSweepArgs& a, FwCtx<N, INREG, SECOND, G>& c,
                                             int r) {
  const int st = r % KQ_RING;
  mbar_expect_tx(&c.mbar[st], SECOND ? 2 * c.row_bytes : c.row_bytes);
  bulk_g2s(c.ring + st * c.stage_cplx, a.X + ((size_t)r * c.K + c.k0) * N, c.row_bytes,
           &c.mbar[st]);
  if (SECOND)
    bulk_g2s(c.ring0 + st * c.stage_cplx, a.Phi0 + ((size_t)r * c.K + c.k0) * N, c.row_bytes,
             &c.mbar[st]);
}

// Run consecutive steps while their planned Taylor degree is MT (0 = generic).
template <int N, bool INREG, bool SECOND, typename G, int MT>
__device__ __forceinline__ int fw_run(const KqSweepArgs& a, FwCtx<N, INREG, SECOND, G>& c, int j,
                                      int len) {
  constexpr int NN = N * N;
  while (j < len && c.splan[j] == MT) {
    const int n = c.base + j;
    const int par = n & 1;
    const double dt_cur = c.sdt[j], g_cur = c.sg[j], sl = c.ssl[j];
    // ---- Im <chi| mu |phi> ||chi|| (+ 0.5 sigma Im <dphi| mu |phi>) -------
    double val;
    if (SECOND) {
      double v1 = 0.0, v2 = 0.0;
#pragma unroll
      for (int r = 0; r < N; ++r) {
        cplx w = c_zero();
#pragma unroll
        for (int cc = 0; cc < N; ++cc) w = c_fma(c.mu[cc * N + r], c.phi[cc], w);
        v1 += c_im_conj_mul(c.chi[r], w);
        v2 += c_im_conj_mul(c.dphi[r], w);
      }
      val = fma(0.5 * c.ssig[j], c.valid ? v2 : 0.0, v1 * c.cnorm);
    } else {
      double e0 = 0.0, e1 = 0.0;
#pragma unroll
      for (int cc = 0; cc < N; ++cc) {
        if (cc & 1)
          e1 += c_im_conj_mul(c.eta[cc], c.phi[cc]);
        else
          e0 += c_im_conj_mul(c.eta[cc], c.phi[cc]);
      }
      val = e0 + e1;
    }
    val = warp_allreduce_sum(val);
    if (c.lane == 0) c.red[par * 32 + c.warp] = val;
    __syncthreads();
    if (threadIdx.x == 0 && n + KQ_RING <= c.last_row) fw_issue_row(a, c, n + KQ_RING);
    double d1;
    if (!c.multi) {
      const double* rp = c.red + par * 32;
      if (c.nwarps <= 4) {
        const double2 p0 = *reinterpret_cast<const double2*>(rp);
        const double2 p1 = *reinterpret_cast<const double2*>(rp + 2);
        d1 = (p0.x + (c.nwarps > 1 ? p0.y : 0.0)) +
             ((c.nwarps > 2 ? p1.x : 0.0) + (c.nwarps > 3 ? p1.y : 0.0));
      } else {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int w = 0;
        for (; w + 3 < c.nwarps; w += 4) {
          s0 += rp[w];
          s1 += rp[w + 1];
          s2 += rp[w + 2];
          s3 += rp[w + 3];
        }
        for (; w < c.nwarps; ++w) s0 += rp[w];
        d1 = (s0 + s1) + (s2 + s3);
      }
    } else {
      const uint32_t tag = a.tag_base + (uint32_t)n + 1u;
      if (c.warp == 0) {
        double acc = (c.lane < c.nwarps) ? c.red[par * 32 + c.lane] : 0.0;
        acc = warp_allreduce_sum(acc);
        if (c.nblk > 1) {
          if (c.lane == 0) slot_store(&a.slots[(size_t)par * c.nblk + blockIdx.x], acc, tag);
          double g2 = 0.0;
          for (int q = c.lane; q < c.nblk; q += 32)
            g2 += slot_wait(&a.slots[(size_t)par * c.nblk + q], tag, c.failed);
          acc = warp_allreduce_sum(g2);
        }
        if (a.world > 1) {
          KqSlot* mine = a.peer_slots[a.rank];
          const size_t goff = KQ_RANK_SLOT_OFFSET;
          if (blockIdx.x == 0 && c.lane < a.world)
            slot_store(a.peer_slots[c.lane] + goff + ((size_t)par * a.world + a.rank) * KQ_LMAX,
                       acc, tag);
          double g3 = 0.0;
          if (c.lane == 0) {
            for (int r = 0; r < a.world; ++r)
              g3 += slot_wait(mine + goff + ((size_t)par * a.world + r) * KQ_LMAX, tag, c.failed);
          }
          acc = __shfl_sync(0xffffffffu, g3, 0);
        }
        if (c.lane == 0) c.tot[par] = acc;
      }
      __syncthreads();
      d1 = c.tot[par];
    }
    // ---- pulse update (optimize.py:471-477) --------------------------------
    const double eps_new = __dadd_rn(g_cur, __dmul_rn(sl, d1));
    const double eps = c.driven ? eps_new : c.c1_fixed;
    if (c.writer) {
      a.opt_pulses[n] = eps_new;
      c.ga = __dadd_rn(c.ga, __dmul_rn(__dmul_rn(sl, __dmul_rn(d1, d1)), dt_cur));
    }
    // ---- forward step under the updated pulse -------------------------------
    G At[NN];
    cplx out[N];
    // the plan was made from the guess pulse with CTA-wide norm bounds; the
    // updated pulse (CTA-uniform) may push the bound out of its binade (rare)
    const double x_new = dt_cur * (fma(fabs(eps_new), c.O1, c.O0) + c.Oc);
    if (MT > 0 && x_new <= c.sbound[j]) {
      c.T.assemble(dt_cur, dt_cur * eps, At);
      step_fixed<N, (MT > 0 ? MT : 1), G>(At, c.phi, out);
    } else {
      int s, m;
      double bound;
      plan_bound(x_new, s, m, bound);
      const double h = dt_cur / (double)s;
      c.T.assemble(h, h * eps, At);
      expmv_generic<N, G>(At, c.phi, out, s, m);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) c.phi[i] = out[i];
    // ---- next step's backward state (TMA ring) ------------------------------
    if (n + 1 <= c.last_row) {
      const int r = n + 1, st = r % KQ_RING;
      mbar_wait(&c.mbar[st], (uint32_t)((r / KQ_RING) & 1));
#pragma unroll
      for (int i = 0; i < N; ++i) {
        c.chi[i] = c.ring[st * c.stage_cplx + (size_t)c.tl * N + i];
        if (SECOND)
          c.dphi[i] = c_sub(c.phi[i], c.ring0[st * c.stage_cplx + (size_t)c.tl * N + i]);
      }
      if (!SECOND) fw_make_eta(c);
    }
    if (SECOND && a.store && c.valid) {
#pragma unroll
      for (int i = 0; i < N; ++i) a.store[((size_t)(n + 1) * c.K + c.k) * N + i] = c.phi[i];
    }
    ++j;
  }
  return j;
}

template <int N, int FSEL, bool SECOND, int BTMAX, typename G>
__global__ void __launch_bounds__(BTMAX, 1) k_fwupd_spec(const KqSweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NN = N * N;
  constexpr bool INREG = (N <= 3);
  const int BT = blockDim.x, tid = threadIdx.x;
  const int K = a.K, NT = a.NT;
  // fall-back of the time-parallel sweep: run only if it asked for it
  if (a.cond_epoch && *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch) return;
  FwCtx<N, INREG, SECOND, G> c;
  c.K = K;
  c.lane = tid & 31;
  c.warp = tid >> 5;
  c.nwarps = (BT + 31) >> 5;
  c.k0 = blockIdx.x * BT;
  int k = c.k0 + tid;
  c.valid = k < K;
  if (!c.valid) k = K - 1;
  c.k = k;
  const int kcta = min(BT, K - c.k0);   // objectives held by this CTA
  c.tl = c.valid ? tid : kcta - 1;      // row slot read by this thread
  c.nblk = gridDim.x;
  c.multi = (c.nblk > 1) || (a.world > 1);
  c.writer = (blockIdx.x == 0 && tid == 0);
  c.failed = false;
  c.ga = 0.0;

  c.red = reinterpret_cast<double*>(smem_raw);             // [2][32]
  c.tot = c.red + 64;                                      // [2] (+6 pad)
  c.mbar = reinterpret_cast<uint64_t*>(c.red + 72);        // [KQ_RING]
  double* sdt = reinterpret_cast<double*>(c.mbar + KQ_RING);
  double* sg = sdt + KQ_NTC;
  double* ssl = sg + KQ_NTC;
  double* ssig = ssl + KQ_NTC;
  double* sbound = ssig + KQ_NTC;
  unsigned char* splan = reinterpret_cast<unsigned char*>(sbound + KQ_NTC);
  c.sdt = sdt;
  c.sg = sg;
  c.ssl = ssl;
  c.ssig = ssig;
  c.sbound = sbound;
  c.splan = splan;
  c.ring = reinterpret_cast<cplx*>(splan + KQ_NTC);
  c.stage_cplx = (size_t)BT * N;
  c.ring0 = c.ring + (size_t)KQ_RING * c.stage_cplx;
  cplx* sm_terms = c.ring0 + (SECOND ? (size_t)KQ_RING * c.stage_cplx : 0);
  c.row_bytes = (uint32_t)kcta * N * sizeof(cplx);
  c.last_row = NT - 1;   // chi(t_r) for steps r = 0..NT-1; Phi0 row r travels with it

  if (tid == 0) {
    for (int st = 0; st < KQ_RING; ++st) mbar_init(&c.mbar[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int r = 0; r < KQ_RING && r <= c.last_row; ++r) fw_issue_row(a, c, r);
  }
  c.T.template load<FSEL>(a.ops + ((size_t)k * 2 + 0) * NN, a.ops + ((size_t)k * 2 + 1) * NN,
                          sm_terms, BT, tid);
  c.opn0 = a.op_norm[k * 2 + 0];
  c.opn1 = a.op_norm[k * 2 + 1];
  c.driven = a.term2pulse[k * 2 + 1] == 0;  // else drift-like or padding
  c.c1_fixed = (a.term2pulse[k * 2 + 1] == -1) ? 1.0 : 0.0;
  const double lam = a.lambda_a[0];
#pragma unroll
  for (int e = 0; e < NN; ++e) c.mu[e] = a.mu[(size_t)k * NN + e];
  c.cnorm = c.valid ? a.chi_norms[k] : 0.0;
  c.O0 = block_max(c.opn0, c.red);
  c.O1 = block_max(c.driven ? c.opn1 : 0.0, c.red);
  c.Oc = block_max(c.driven ? 0.0 : c.c1_fixed * c.opn1, c.red);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    c.phi[i] = a.state0[(size_t)k * N + i];
    c.dphi[i] = c_zero();
  }
  mbar_wait(&c.mbar[0], 0);   // row 0
#pragma unroll
  for (int i = 0; i < N; ++i) c.chi[i] = c.ring[(size_t)c.tl * N + i];
  fw_make_eta(c);
  if (SECOND && a.store && c.valid) {
#pragma unroll
    for (int i = 0; i < N; ++i) a.store[((size_t)0 * K + k) * N + i] = c.phi[i];
  }

  for (int base = 0; base < NT; base += KQ_NTC) {
    const int len = min(KQ_NTC, NT - base);
    // stage the per-step scalars and the Taylor plan of this chunk
    __syncthreads();
    for (int i = tid; i < len; i += BT) {
      const double dti = a.dt[base + i], gi = a.pulses[base + i];
      sdt[i] = dti;
      sg[i] = gi;
      ssl[i] = a.shape[base + i] / lam;   // S/lambda as in optimize.py:474
      if (SECOND) ssig[i] = a.sigma[base + i];
      int s, m;
      double bound;
      plan_bound(dti * (fma(fabs(gi), c.O1, c.O0) + c.Oc), s, m, bound);
      sbound[i] = bound;
      splan[i] = (s == 1 && m <= KQ_MFIX) ? (unsigned char)m : (unsigned char)0;
    }
    __syncthreads();
    c.base = base;
    int j = 0;
    while (j < len) {
      switch (splan[j]) {
        case 1: j = fw_run<N, INREG, SECOND, G, 1>(a, c, j, len); break;
        case 2: j = fw_run<N, INREG, SECOND, G, 2>(a, c, j, len); break;
        case 3: j = fw_run<N, INREG, SECOND, G, 3>(a, c, j, len); break;
        case 4: j = fw_run<N, INREG, SECOND, G, 4>(a, c, j, len); break;
        case 5: j = fw_run<N, INREG, SECOND, G, 5>(a, c, j, len); break;
        case 6: j = fw_run<N, INREG, SECOND, G, 6>(a, c, j, len); break;
        case 7: j = fw_run<N, INREG, SECOND, G, 7>(a, c, j, len); break;
        case 8: j = fw_run<N, INREG, SECOND, G, 8>(a, c, j, len); break;
        default: j = fw_run<N, INREG, SECOND, G, 0>(a, c, j, len); break;
      }
    }
  }
  if (a.stateT && c.valid) {
#pragma unroll
    for (int i = 0; i < N; ++i) a.stateT[(size_t)k * N + i] = c.phi[i];
  }
  if (c.writer) a.g_a[0] = c.ga;
  if (c.failed) atomicExch(a.status, (int)-4);
}
