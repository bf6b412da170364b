// Time-parallel fused pulse update + forward sweep (optimize.py:449-500 of the
// reference) for the specialised problem shape (N <= 4, one drift + one control
// term, one pulse).
//
// The sequential sweep is a chain of nt-1 dependent steps:
//     eps[n] = guess[n] + (S[n]/lambda) Im sum_k <chi_k[n]| mu |phi_k[n]>,
//     phi_k[n+1] = U_k(eps[n]) phi_k[n].
// Because phi_k[n] depends on eps[0..n-1] only, the updated pulse is the unique
// fixed point of the *causal* map  eps -> guess + (S/lambda) F(eps), where F
// propagates all objectives under eps and evaluates the overlaps at every time
// step.  One evaluation of F is fully parallel in time; the Picard iteration
// eps_{j+1} = guess + (S/lambda) F(eps_j) converges like (C T)^j / j!  (Volterra
// structure; C4: 9 iterations to 1e-15) and is exact after at most nt-1
// iterations whatever the coupling.  Iterating to |eps_{j+1} - eps_j| <= rtol
// max|eps| reproduces the sequential result to rounding.
//
// Mapping: CTA = Q objectives x TC time chunks of W steps (thread = one chunk of
// one objective).  One evaluation of F:
//   pass A  every thread propagates the N basis vectors through its chunk
//           (chunk propagator M_t, registers);
//   scan    Kogge-Stone inclusive scan of M_t over the lanes (N x N complex
//           products, SHFL), warp totals through shared memory -> state at the
//           start of every chunk;
//   pass B  every thread propagates that state through its chunk and evaluates
//           Im <mu^dag chi ||chi|| | phi> at each step (mu^dag chi prepared once
//           per launch in shared memory);
//   sum     over the objectives: inside the CTA in shared memory, across CTAs
//           through flag-tagged 16-byte slots in global memory (kq_common.cuh):
//           every CTA publishes its partial sums for all time steps, CTA c
//           reduces time slice c in a fixed order and publishes the updated
//           pulse values, every CTA reads the whole updated pulse.  Two
//           one-way L2 hops per iteration, no grid barrier, deterministic.
// A final evaluation under the converged pulse stores phi(T) (and all forward
// states for second order).  If the iteration does not converge in pic_maxit
// rounds the kernel requests the sequential kernel through status[1].
#pragma once
#include "kq_spec.cuh"

#define KQ_PIC_WPO 8      // warps per objective at most (TC <= 256)
#define KQ_PIC_BT 256     // threads per CTA the kernels are compiled for

__device__ __forceinline__ cplx shfl_up_c(cplx v, int d) {
  return make_double2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}

// C = A * B (column-major N x N)
template <int N>
__device__ __forceinline__ void mat_mul(const cplx (&A)[N * N], const cplx (&B)[N * N],
                                        cplx (&C)[N * N]) {
#pragma unroll
  for (int c = 0; c < N; ++c)
#pragma unroll
    for (int r = 0; r < N; ++r) {
      cplx a0 = c_zero(), a1 = c_zero();
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (k & 1)
          a1 = c_fma(A[k * N + r], B[c * N + k], a1);
        else
          a0 = c_fma(A[k * N + r], B[c * N + k], a0);
      }
      C[c * N + r] = (N > 1) ? c_add(a0, a1) : a0;
    }
}

// y = A * x
template <int N>
__device__ __forceinline__ void mat_vec(const cplx* A, const cplx (&x)[N], cplx (&y)[N]) {
#pragma unroll
  for (int r = 0; r < N; ++r) {
    cplx a0 = c_zero(), a1 = c_zero();
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (k & 1)
        a1 = c_fma(A[k * N + r], x[k], a1);
      else
        a0 = c_fma(A[k * N + r], x[k], a0);
    }
    y[r] = (N > 1) ? c_add(a0, a1) : a0;
  }
}

// ---- one propagation step, prepared once and applied to several vectors ------
template <int N, bool INREG, typename G>
struct StepOp {
  G At[N * N];
  int s, m;
  __device__ __forceinline__ void prepare(const SpecTerms<N, INREG, G>& T, double dt, double eps,
                                          int s_, int m_) {
    s = s_;
    m = m_;
    const double h = (s_ == 1) ? dt : dt / (double)s_;
    T.assemble(h, h * eps, At);
  }
  __device__ __forceinline__ void apply(cplx (&y)[N]) const {
    cplx out[N];
    expmv_generic<N, G>(At, y, out, s, m);
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = out[i];
  }
};

template <int P>
__device__ __forceinline__ void pic_poly(double z, double& C, double& S) {
  C = kq_inv_fact(2 * (P - 1));
  S = kq_inv_fact(2 * (P - 1) + 1);
#pragma unroll
  for (int j = P - 2; j >= 0; --j) {
    C = fma(-z, C, kq_inv_fact(2 * j));
    S = fma(-z, S, kq_inv_fact(2 * j + 1));
  }
}

// N = 2 with a real generator: closed form (see expmv2_real in kq_spec.cuh);
// the polynomials C(z), S(z) are evaluated once per step and shared by all
// vectors the step is applied to.
template <bool INREG>
struct StepOp<2, INREG, double> {
  double C, S, dl, b, c, st, ct;
  int s;
  bool phase;
  __device__ __forceinline__ void prepare(const SpecTerms<2, INREG, double>& T, double dt,
                                          double eps, int s_, int m_) {
    s = s_;
    const double h = (s_ == 1) ? dt : dt / (double)s_;
    double R[4];
    T.assemble(h, h * eps, R);
    const double a = R[0], d = R[3];
    c = R[1];
    b = R[2];
    const double t = 0.5 * (a + d);
    dl = 0.5 * (a - d);
    const double z = fma(dl, dl, b * c);
    const int P = max(3, (m_ + 4) >> 1);   // 2P >= m + 3
    switch (P) {
      case 3: pic_poly<3>(z, C, S); break;
      case 4: pic_poly<4>(z, C, S); break;
      case 5: pic_poly<5>(z, C, S); break;
      case 6: pic_poly<6>(z, C, S); break;
      default: {
        C = c_kq_tables.invfact[2 * (P - 1)];
        S = c_kq_tables.invfact[2 * (P - 1) + 1];
        for (int j = P - 2; j >= 0; --j) {
          C = fma(-z, C, c_kq_tables.invfact[2 * j]);
          S = fma(-z, S, c_kq_tables.invfact[2 * j + 1]);
        }
      }
    }
    phase = (t != 0.0);
    st = 0.0;
    ct = 1.0;
    if (phase) sincos(t, &st, &ct);
  }
  __device__ __forceinline__ void apply(cplx (&y)[2]) const {
    for (int rep = 0; rep < s; ++rep) {
      const cplx v0 = y[0], v1 = y[1];
      const cplx w0 = make_double2(fma(dl, v0.x, b * v1.x), fma(dl, v0.y, b * v1.y));
      const cplx w1 = make_double2(fma(-dl, v1.x, c * v0.x), fma(-dl, v1.y, c * v0.y));
      cplx u0 = make_double2(fma(-S, w0.y, C * v0.x), fma(S, w0.x, C * v0.y));
      cplx u1 = make_double2(fma(-S, w1.y, C * v1.x), fma(S, w1.x, C * v1.y));
      if (phase) {
        u0 = make_double2(fma(-st, u0.y, ct * u0.x), fma(st, u0.x, ct * u0.y));
        u1 = make_double2(fma(-st, u1.y, ct * u1.x), fma(st, u1.x, ct * u1.y));
      }
      y[0] = u0;
      y[1] = u1;
    }
  }
};

// CTA-wide maxima of four non-negative doubles (all threads call).
// scratch: [4][32] doubles.  Two barriers.
__device__ __forceinline__ void block_max4(double& v0, double& v1, double& v2, double& v3,
                                           double* scratch) {
  v0 = warp_allreduce_max(v0);
  v1 = warp_allreduce_max(v1);
  v2 = warp_allreduce_max(v2);
  v3 = warp_allreduce_max(v3);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    scratch[warp] = v0;
    scratch[32 + warp] = v1;
    scratch[64 + warp] = v2;
    scratch[96 + warp] = v3;
  }
  __syncthreads();
  double r0 = scratch[0], r1 = scratch[32], r2 = scratch[64], r3 = scratch[96];
  for (int w = 1; w < nw; ++w) {
    r0 = fmax(r0, scratch[w]);
    r1 = fmax(r1, scratch[32 + w]);
    r2 = fmax(r2, scratch[64 + w]);
    r3 = fmax(r3, scratch[96 + w]);
  }
  __syncthreads();
  v0 = r0;
  v1 = r1;
  v2 = r2;
  v3 = r3;
}

// CTA-wide sum in a fixed order (warp butterfly, then the warps in order).
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  v = warp_allreduce_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = scratch[0];
  for (int w = 1; w < nw; ++w) r += scratch[w];
  __syncthreads();
  return r;
}

// Shared state of one thread of the time-parallel kernels.
template <int N, bool INREG, typename G>
struct PicCtx {
  SpecTerms<N, INREG, G> T;
  cplx start[N];     // state of this objective at the start of the sweep
  int q, t, lane, wq, k, W, TC, NT;
  bool driven;
  double c1_fixed;
  const double* dt;
  const double* seps;   // [W][TC] transposed pulse values of the current iterate
  cplx* wtot;           // [Q][KQ_PIC_WPO][N*N] warp totals
};

// Pass A + scan: state at the start of this thread's chunk under the pulse in
// c.seps.  Contains one __syncthreads (all threads of the CTA must call).
template <int N, bool INREG, typename G>
__device__ __forceinline__ void pic_chunk_start(const PicCtx<N, INREG, G>& c, int s, int m,
                                                cplx (&b)[N]) {
  constexpr int NN = N * N;
  cplx M[NN];
#pragma unroll
  for (int e = 0; e < NN; ++e) M[e] = c_make((e % N) == (e / N) ? 1.0 : 0.0, 0.0);
  for (int w = 0; w < c.W; ++w) {
    const int n = c.t * c.W + w;
    if (n < c.NT) {
      const double eps = c.driven ? c.seps[w * c.TC + c.t] : c.c1_fixed;
      StepOp<N, INREG, G> op;
      op.prepare(c.T, c.dt[n], eps, s, m);
#pragma unroll
      for (int v = 0; v < N; ++v) {
        cplx y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = M[v * N + i];
        op.apply(y);
#pragma unroll
        for (int i = 0; i < N; ++i) M[v * N + i] = y[i];
      }
    }
  }
  // inclusive scan over the lanes: M <- M_lane * ... * M_0
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    cplx O[NN];
#pragma unroll
    for (int e = 0; e < NN; ++e) O[e] = shfl_up_c(M[e], off);
    if (c.lane >= off) {
      cplx Cm[NN];
      mat_mul<N>(M, O, Cm);
#pragma unroll
      for (int e = 0; e < NN; ++e) M[e] = Cm[e];
    }
  }
  cplx* wt = c.wtot + (size_t)c.q * KQ_PIC_WPO * NN;
  if (c.lane == 31) {
#pragma unroll
    for (int e = 0; e < NN; ++e) wt[c.wq * NN + e] = M[e];
  }
  __syncthreads();
  cplx v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = c.start[i];
  for (int w = 0; w < c.wq; ++w) {
    cplx o[N];
    mat_vec<N>(wt + w * NN, v, o);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = o[i];
  }
  cplx e_[N];
  mat_vec<N>(M, v, e_);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const cplx up = shfl_up_c(e_[i], 1);
    b[i] = (c.lane == 0) ? v[i] : up;
  }
}

// Taylor plan for the whole CTA from the largest scaled norm of any step.
__device__ __forceinline__ void pic_plan(double xmax, int& s, int& m) {
  double bound;
  plan_bound(xmax, s, m, bound);
}

// shared: [scratch 128][seps NTP][dsm Q*NTP][eta Q*NTP*N cplx][wtot Q*WPO*NN cplx][terms]
template <int N, int FSEL, bool SECOND, typename G>
__global__ void __launch_bounds__(KQ_PIC_BT, 1) k_fwupd_picard(const KqSweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NN = N * N;
  constexpr bool INREG = (N <= 3);
  const int tid = threadIdx.x, BT = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = BT >> 5;
  const int Q = a.pic_Q, TC = a.pic_TC, W = a.pic_W, NT = a.NT, K = a.K;
  const int NTP = TC * W;
  const int nblk = gridDim.x;
  const bool single = (nblk == 1);

  double* scratch = reinterpret_cast<double*>(smem_raw);   // [128]
  double* seps = scratch + 128;                             // [NTP]  index w*TC + t
  double* dsm = seps + NTP;                                 // [Q][NTP]
  cplx* eta = reinterpret_cast<cplx*>(dsm + (size_t)Q * NTP);   // [Q][W][N][TC]
  cplx* wtot = eta + (size_t)Q * NTP * N;                   // [Q][WPO][NN]
  G* sterms = reinterpret_cast<G*>(wtot + (size_t)Q * KQ_PIC_WPO * NN);   // [Q][2][NN] (N = 4)

  PicCtx<N, INREG, G> c;
  c.q = tid / TC;
  c.t = tid - c.q * TC;
  c.lane = lane;
  c.wq = c.t >> 5;
  c.W = W;
  c.TC = TC;
  c.NT = NT;
  c.dt = a.dt;
  c.seps = seps;
  c.wtot = wtot;
  int k = blockIdx.x * Q + c.q;
  const bool valid = k < K;
  if (!valid) k = K - 1;
  c.k = k;
  if (INREG) {
    c.T.template load<FSEL>(a.ops + ((size_t)k * 2 + 0) * NN, a.ops + ((size_t)k * 2 + 1) * NN,
                            nullptr, BT, tid);
  } else {
    if (c.t < NN) {
      sterms[(c.q * 2 + 0) * NN + c.t] = g_load<FSEL>(a.ops[((size_t)k * 2 + 0) * NN + c.t], G());
      sterms[(c.q * 2 + 1) * NN + c.t] = g_load<FSEL>(a.ops[((size_t)k * 2 + 1) * NN + c.t], G());
    }
    c.T.s0 = sterms + (size_t)c.q * 2 * NN;
    c.T.s1 = c.T.s0 + NN;
    c.T.stride = 1;
  }
  const double opn0 = a.op_norm[k * 2 + 0], opn1 = a.op_norm[k * 2 + 1];
  c.driven = a.term2pulse[k * 2 + 1] == 0;
  c.c1_fixed = (a.term2pulse[k * 2 + 1] == -1) ? 1.0 : 0.0;
  const double lam = a.lambda_a[0];
  cplx mu[NN];
#pragma unroll
  for (int e = 0; e < NN; ++e) mu[e] = a.mu[(size_t)k * NN + e];
  const double cnorm = valid ? a.chi_norms[k] : 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) c.start[i] = a.state0[(size_t)k * N + i];

  // eta[n] = mu^dag chi[n] ||chi||  for the steps of this chunk
  for (int w = 0; w < W; ++w) {
    const int n = c.t * W + w;
    cplx chi[N];
#pragma unroll
    for (int i = 0; i < N; ++i)
      chi[i] = (n < NT) ? a.X[((size_t)n * K + k) * N + i] : c_zero();
#pragma unroll
    for (int cc = 0; cc < N; ++cc) {
      cplx acc = c_zero();
#pragma unroll
      for (int r = 0; r < N; ++r) acc = c_fma_conj(mu[cc * N + r], chi[r], acc);
      eta[(((size_t)c.q * W + w) * N + cc) * TC + c.t] =
          make_double2(acc.x * cnorm, acc.y * cnorm);
    }
  }
  // CTA-wide operator norm bounds and the first iterate (the guess pulse)
  double O0 = opn0, O1 = c.driven ? opn1 : 0.0, Oc = c.driven ? 0.0 : c.c1_fixed * opn1;
  double dummy = 0.0;
  block_max4(O0, O1, Oc, dummy, scratch);
  double xm = 0.0, em = 0.0, dm = 0.0, bad = 0.0;
  for (int n = tid; n < NTP; n += BT) {
    const double e = (n < NT) ? a.pulses[n] : 0.0;
    seps[(n % W) * TC + n / W] = e;
    if (n < NT) xm = fmax(xm, a.dt[n] * (fma(fabs(e), O1, O0) + Oc));
  }
  block_max4(xm, em, dm, bad, scratch);   // also orders the smem writes above

  const int Wc = (NT + nblk - 1) / nblk;   // time slice reduced by each CTA
  const int n_lo = blockIdx.x * Wc;
  bool failed = false, converged = false;
  double ga_acc = 0.0;   // single: this thread's share; multi: lane 0 of each warp
  int it = 0;
  while (true) {
    ++it;
    int s, m;
    pic_plan(xm, s, m);
    cplx y[N];
    pic_chunk_start<N, INREG, G>(c, s, m, y);
    // ---- pass B: overlaps at every step of the chunk -------------------------
    for (int w = 0; w < W; ++w) {
      const int n = c.t * W + w;
      double d = 0.0;
      if (n < NT) {
        double e0 = 0.0, e1 = 0.0;
#pragma unroll
        for (int cc = 0; cc < N; ++cc) {
          const cplx et = eta[(((size_t)c.q * W + w) * N + cc) * TC + c.t];
          if (cc & 1)
            e1 += c_im_conj_mul(et, y[cc]);
          else
            e0 += c_im_conj_mul(et, y[cc]);
        }
        d = e0 + e1;
        if (SECOND) {
          if (n > 0 && valid) {
            double v2 = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r) {
              cplx wv = c_zero();
#pragma unroll
              for (int cc = 0; cc < N; ++cc) wv = c_fma(mu[cc * N + r], y[cc], wv);
              const cplx dphi = c_sub(y[r], a.Phi0[((size_t)n * K + k) * N + r]);
              v2 += c_im_conj_mul(dphi, wv);
            }
            d = fma(0.5 * a.sigma[n], v2, d);
          }
        }
        const double eps = c.driven ? seps[w * TC + c.t] : c.c1_fixed;
        StepOp<N, INREG, G> op;
        op.prepare(c.T, a.dt[n], eps, s, m);
        op.apply(y);
      }
      dsm[(size_t)c.q * NTP + w * TC + c.t] = d;
    }
    __syncthreads();
    // ---- sum over the objectives, pulse update --------------------------------
    xm = 0.0;
    em = 0.0;
    dm = 0.0;
    ga_acc = 0.0;
    if (single) {
      for (int n = tid; n < NT; n += BT) {
        const int idx = (n % W) * TC + n / W;
        double d1 = dsm[idx];
        for (int qq = 1; qq < Q; ++qq) d1 += dsm[(size_t)qq * NTP + idx];
        const double sl = a.shape[n] / lam;
        const double dtn = a.dt[n];
        const double e_new = __dadd_rn(a.pulses[n], __dmul_rn(sl, d1));
        ga_acc = __dadd_rn(ga_acc, __dmul_rn(__dmul_rn(sl, __dmul_rn(d1, d1)), dtn));
        dm = fmax(dm, fabs(e_new - seps[idx]));
        em = fmax(em, fabs(e_new));
        xm = fmax(xm, dtn * (fma(fabs(e_new), O1, O0) + Oc));
        if (!(fabs(e_new) < 1.0e150)) bad = 1.0;
        seps[idx] = e_new;
      }
    } else {
      const uint32_t tag = a.tag_base + (uint32_t)it;
      const size_t stride = (size_t)a.pic_stride;
      // stage 1: this CTA's partial sums for every time step
      for (int n = tid; n < NT; n += BT) {
        const int idx = (n % W) * TC + n / W;
        double d1 = dsm[idx];
        for (int qq = 1; qq < Q; ++qq) d1 += dsm[(size_t)qq * NTP + idx];
        slot_store(&a.pic_part[(size_t)blockIdx.x * stride + n], d1, tag);
      }
      // stage 2: reduce time slice [n_lo, n_lo + Wc) over all CTAs (fixed order)
      for (int ni = warp; ni < Wc; ni += nwarps) {
        const int n = n_lo + ni;
        if (n < NT) {
          double acc = 0.0;
          for (int cb = lane; cb < nblk; cb += 32)
            acc += slot_wait(&a.pic_part[(size_t)cb * stride + n], tag, failed);
          acc = warp_allreduce_sum(acc);
          if (lane == 0) {
            const double sl = a.shape[n] / lam;
            const double e_new = __dadd_rn(a.pulses[n], __dmul_rn(sl, acc));
            ga_acc = __dadd_rn(ga_acc, __dmul_rn(__dmul_rn(sl, __dmul_rn(acc, acc)), a.dt[n]));
            slot_store(&a.pic_eps[n], e_new, tag);
          }
        }
      }
      // stage 3: the whole updated pulse
      for (int n = tid; n < NT; n += BT) {
        const int idx = (n % W) * TC + n / W;
        const double e_new = slot_wait(&a.pic_eps[n], tag, failed);
        dm = fmax(dm, fabs(e_new - seps[idx]));
        em = fmax(em, fabs(e_new));
        xm = fmax(xm, a.dt[n] * (fma(fabs(e_new), O1, O0) + Oc));
        if (!(fabs(e_new) < 1.0e150)) bad = 1.0;
        seps[idx] = e_new;
      }
      if (failed) bad = 2.0;
    }
    block_max4(xm, em, dm, bad, scratch);
    if (bad > 0.0) break;
    if (dm <= a.pic_rtol * em) {
      converged = true;
      break;
    }
    if (it >= a.pic_maxit) break;
  }

  if (!converged) {
    // ask for the sequential kernel (launched right after this one)
    if (blockIdx.x == 0 && tid == 0) {
      a.status[1] = (int)a.epoch;
      if (bad > 1.5) atomicExch(a.status, (int)-4);
    }
    return;
  }
  // ---- outputs ----------------------------------------------------------------
  if (blockIdx.x == 0) {
    for (int n = tid; n < NT; n += BT) a.opt_pulses[n] = seps[(n % W) * TC + n / W];
  }
  if (single) {
    const double ga = block_sum(ga_acc, scratch);
    if (tid == 0) a.g_a[0] = ga;
  } else {
    const uint32_t tagf = a.tag_base + (uint32_t)a.pic_maxit + 1u;
    double ga = block_sum((lane == 0) ? ga_acc : 0.0, scratch);
    if (tid == 0) slot_store(&a.pic_ga[blockIdx.x], ga, tagf);
    if (blockIdx.x == 0 && warp == 0) {
      double acc = 0.0;
      for (int cb = lane; cb < nblk; cb += 32) acc += slot_wait(&a.pic_ga[cb], tagf, failed);
      acc = warp_allreduce_sum(acc);
      if (lane == 0) {
        a.g_a[0] = acc;
        if (failed) atomicExch(a.status, (int)-4);
      }
    }
  }
  if (tid == 0 && blockIdx.x == 0) a.status[2] = it;   // Picard iterations used (diagnostics)
  // final evaluation under the converged pulse: phi(T), forward states
  {
    int s, m;
    pic_plan(xm, s, m);
    cplx y[N];
    pic_chunk_start<N, INREG, G>(c, s, m, y);
    const bool store = SECOND && a.store && valid;
    if (store && c.t == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) a.store[(size_t)k * N + i] = y[i];
    }
    for (int w = 0; w < W; ++w) {
      const int n = c.t * W + w;
      if (n < NT) {
        const double eps = c.driven ? seps[w * TC + c.t] : c.c1_fixed;
        StepOp<N, INREG, G> op;
        op.prepare(c.T, a.dt[n], eps, s, m);
        op.apply(y);
        if (store) {
#pragma unroll
          for (int i = 0; i < N; ++i) a.store[((size_t)(n + 1) * K + k) * N + i] = y[i];
        }
        if (n == NT - 1 && a.stateT && valid) {
#pragma unroll
          for (int i = 0; i < N; ++i) a.stateT[(size_t)k * N + i] = y[i];
        }
      }
    }
  }
}
