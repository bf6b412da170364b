// Update / forward sweep (optimize.py:449-500 of the reference) for FEW objectives with
// SEVERAL controls or terms and small state vectors: the Lambda systems of the reference's
// notebooks 02 / 03 / 08 (N = 3, four real controls, one to five objectives).
//
// The sweep is a chain of NT dependent steps; with so few objectives nothing but the length
// of one link counts.  The thread-per-objective kernel (kq_small.cuh) issues the whole step
// -- L overlaps, the assembly of M terms, m Taylor terms of an N x N product -- from one
// thread: about 1300 FP64 instructions per step behind each other, 4.3 us per step measured.
// Here ONE warp runs the chain with lane = (objective k, row r):
//   * a parallel pre-pass (k_lanes_prep, one thread per (time step, lane)) turns the stored
//     backward states into zeta_l[n][k][c] = ||chi_k|| sum_r conj(chi_k[n][r]) mu_lk[r, c], so
//     that the overlap Im <chi_k[n]| mu_lk |phi_k> = Im sum_c zeta_l[n][k][c] phi_k[c] costs
//     two DFMA per lane and control, and packs the per-step scalars {S_l/lambda_l, guess_l, dt};
//   * the sum over rows and objectives (optimize.py:454-470) is a butterfly over the lanes;
//   * every lane keeps row r of all generator terms in registers (already multiplied by the
//     equation-of-motion factor), assembles its row of A under the UPDATED pulses and carries
//     out the Taylor/Horner recurrence of propagators.expm on its component of the state; the
//     other components come from the neighbouring lanes by shuffles (no shared memory, no
//     barrier anywhere in the kernel);
//   * records of the step after next are loaded while the current step computes.
// First order, one GPU; everything else stays with the kernels of kq_small.cuh.
#pragma once
#include "kq_small.cuh"
#include "kq_lanes_geom.cuh"

// ---- pre-pass: grid ceil(NT / 8), block 256 = 8 time steps x 32 lanes --------------------
__global__ void __launch_bounds__(256) k_lanes_prep(const KqSweepArgs a, const KqLanes d) {
  if (a.cond_epoch && *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch) return;
  const int K = a.K, N = a.N, NT = a.NT, L = a.L, NN = N * N;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= NT) return;
  const int k = lane / d.NP, c = lane - k * d.NP;
  const bool act = k < K && c < N;
  cplx chi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    chi[r] = (act && r < N) ? a.X[((size_t)n * K + k) * N + r] : c_zero();
  const double cn = act ? a.chi_norms[k] : 0.0;
  for (int l = 0; l < L; ++l) {
    cplx z = c_zero();
    if (act) {
      const cplx* mu = a.mu + ((size_t)k * L + l) * NN + (size_t)c * N;   // column c
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r < N) z = c_fma_conj(chi[r], mu[r], z);
    }
    d.zeta[((size_t)n * L + l) * 32 + lane] = c_make(z.x * cn, z.y * cn);
  }
  double* sc = d.scal + (size_t)n * KQ_LN_SC;
  if (lane < KQ_LN_LMAX) {
    const bool in = lane < L;
    sc[lane] = in ? a.shape[(size_t)lane * NT + n] / a.lambda_a[lane] : 0.0;   // optimize.py:474
    sc[KQ_LN_LMAX + lane] = in ? a.pulses[(size_t)lane * NT + n] : 0.0;
  } else if (lane == KQ_LN_LMAX) {
    sc[2 * KQ_LN_LMAX] = a.dt[n];
    sc[2 * KQ_LN_LMAX + 1] = 0.0;
  }
}

// ---- the chain: one warp ------------------------------------------------------------------
struct LnRec {
  cplx z[KQ_LN_LMAX];
  double2 s[KQ_LN_SC / 2];
};
__device__ __forceinline__ void ln_load(LnRec& R, const KqLanes& d, int n, int L, int lane) {
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l)
    R.z[l] = (l < L) ? d.zeta[((size_t)n * L + l) * 32 + lane] : c_zero();
  const double2* sp = reinterpret_cast<const double2*>(d.scal + (size_t)n * KQ_LN_SC);
#pragma unroll
  for (int i = 0; i < KQ_LN_SC / 2; ++i) R.s[i] = sp[i];
}

template <int N>
struct LnState {
  cplx y;                          // this lane's component of phi_k
  double ga[KQ_LN_LMAX];
  cplx T[KQ_MMAX_SMALL][N];        // row r of f * T_m
  int t2p[KQ_MMAX_SMALL];
  double opn[KQ_MMAX_SMALL];
};

template <int N>
__device__ __forceinline__ void ln_step(const KqSweepArgs& a, const KqLanes& d, const LnRec& R,
                                        LnState<N>& S, int n, int L, int M, int lane, int base) {
  const KqTables& T = c_kq_tables;
  const double sl[KQ_LN_LMAX] = {R.s[0].x, R.s[0].y, R.s[1].x, R.s[1].y};
  const double g[KQ_LN_LMAX] = {R.s[2].x, R.s[2].y, R.s[3].x, R.s[3].y};
  const double dtn = R.s[4].x;
  // ---- Im <chi_k| mu_lk |phi_k> ||chi_k||, summed over rows and objectives
  double o[KQ_LN_LMAX];
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) o[l] = fma(R.z[l].x, S.y.y, R.z[l].y * S.y.x);
  for (int off = d.span >> 1; off > 0; off >>= 1) {
#pragma unroll
    for (int l = 0; l < KQ_LN_LMAX; ++l)
      if (l < L) o[l] += __shfl_xor_sync(0xffffffffu, o[l], off);
  }
  // ---- updated pulse values, rounded like optimize.py:474-477
  double eps[KQ_LN_LMAX];
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) {
    eps[l] = __dadd_rn(g[l], __dmul_rn(sl[l], o[l]));
    S.ga[l] = __dadd_rn(S.ga[l], __dmul_rn(__dmul_rn(sl[l], __dmul_rn(o[l], o[l])), dtn));
    if (lane == 0 && l < L) a.opt_pulses[(size_t)l * a.NT + n] = eps[l];
  }
  // ---- this lane's row of f A under the updated pulses
  cplx A[N];
#pragma unroll
  for (int c = 0; c < N; ++c) A[c] = c_zero();
  double x = 0.0;
#pragma unroll
  for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
    if (m < M) {
      const int t = S.t2p[m];
      double coef = (t == -1) ? 1.0 : 0.0;
#pragma unroll
      for (int l = 0; l < KQ_LN_LMAX; ++l) coef = (t == l) ? eps[l] : coef;
      x = fma(fabs(coef), S.opn[m], x);
#pragma unroll
      for (int c = 0; c < N; ++c) A[c] = c_fma_real(coef, S.T[m][c], A[c]);
    }
  }
  x *= dtn;
  // one Taylor plan for the warp: the largest bound of its objectives
  {
    const int hi = __reduce_max_sync(0xffffffffu, __double2hiint(x));
    x = __hiloint2double(hi, (int)0xffffffff);
  }
  int s, mdeg;
  double xs;
  taylor_plan(T, x, s, mdeg, xs);
  const double h = (s == 1) ? dtn : dtn / (double)s;
  for (int rep = 0; rep < s; ++rep) {
    const cplx v = S.y;
    cplx y = v;
    for (int j = mdeg; j >= 1; --j) {
      const double cj = h * T.inv[j];
      cplx w0 = c_zero(), w1 = c_zero();
#pragma unroll
      for (int c = 0; c < N; ++c) {
        cplx yc;
        yc.x = __shfl_sync(0xffffffffu, y.x, base + c);
        yc.y = __shfl_sync(0xffffffffu, y.y, base + c);
        if (c & 1) w1 = c_fma(A[c], yc, w1);
        else w0 = c_fma(A[c], yc, w0);
      }
      const cplx w = (N > 1) ? c_add(w0, w1) : w0;
      y = c_fma_real(cj, w, v);
    }
    S.y = y;
  }
}

template <int N, int FSEL>
__global__ void __launch_bounds__(32, 1) k_fwupd_lanes(const KqSweepArgs a, const KqLanes d) {
  if (a.cond_epoch && *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch) return;
  const int K = a.K, NT = a.NT, M = a.M, L = a.L;
  constexpr int NN = N * N;
  const int lane = threadIdx.x;
  const int k = lane / d.NP, r = lane - k * d.NP;
  const bool act = k < K && r < N;
  const int base = lane - r;
  LnState<N> S;
#pragma unroll
  for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
    S.t2p[m] = -2;
    S.opn[m] = 0.0;
#pragma unroll
    for (int c = 0; c < N; ++c) S.T[m][c] = c_zero();
    if (m < M && act) {
      S.t2p[m] = a.term2pulse[k * M + m];
      S.opn[m] = a.op_norm[k * M + m];
#pragma unroll
      for (int c = 0; c < N; ++c)
        S.T[m][c] = apply_f<FSEL>(a.ops[((size_t)k * M + m) * NN + c * N + r]);
    }
  }
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) S.ga[l] = 0.0;
  S.y = act ? a.state0[(size_t)k * N + r] : c_zero();
  LnRec RA, RB;
  ln_load(RA, d, 0, L, lane);
  ln_load(RB, d, min(1, NT - 1), L, lane);
  for (int n = 0; n < NT; n += 2) {
    {
      const LnRec R = RA;
      ln_load(RA, d, min(n + 2, NT - 1), L, lane);
      ln_step<N>(a, d, R, S, n, L, M, lane, base);
    }
    if (n + 1 < NT) {
      const LnRec R = RB;
      ln_load(RB, d, min(n + 3, NT - 1), L, lane);
      ln_step<N>(a, d, R, S, n + 1, L, M, lane, base);
    }
  }
  if (a.stateT && act) a.stateT[(size_t)k * N + r] = S.y;
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < KQ_LN_LMAX; ++l)
      if (l < L) a.g_a[l] = S.ga[l];
  }
}
