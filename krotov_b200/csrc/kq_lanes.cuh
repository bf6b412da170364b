// Update / forward sweep (optimize.py:449-500 of the reference) for FEW objectives whose
// generators are small or sparse: the Lambda systems of the reference's notebooks 02 / 03 / 08
// (N = 3, four real controls, one to five objectives) and the 17-level transmon of notebook 05
// (tridiagonal drive, diagonal drift).
//
// The sweep is a chain of NT dependent steps; with so few objectives nothing but the length
// of one link counts.  The thread-per-objective kernel (kq_small.cuh) issues the whole step
// -- L overlaps, the assembly of M terms, m Taylor terms of an N x N product -- from one
// thread (4.3 us per step measured for N = 3, L = 4); the lane-per-row kernel (kq_warp.cuh)
// multiplies dense padded rows through shared memory (13 us per step for N = 17).  Here
// lane = (objective k, row r), NP = 2^ceil(log2 N) lanes per objective, and a lane keeps only
// the NZ <= 4 non-zero entries of its row (the union pattern of all terms, plus the diagonal):
//   * a parallel pre-pass (k_rows_prep, one thread per (time step, lane)) turns the stored
//     backward states into zeta_l[n][k][c] = ||chi_k|| sum_r conj(chi_k[n][r]) mu_lk[r, c], so
//     that the overlap Im <chi_k[n]| mu_lk |phi_k> = Im sum_c zeta_l[n][k][c] phi_k[c] costs
//     two DFMA per lane and control, and packs the per-step scalars {S_l/lambda_l, guess_l, dt};
//   * the sum over rows and objectives (optimize.py:454-470) is a butterfly over the lanes
//     (plus one shared-memory hop when the objectives fill several warps);
//   * every lane keeps its entries of all generator terms in registers (already multiplied by
//     the equation-of-motion factor), assembles its row of A under the UPDATED pulses and
//     carries out the Taylor/Horner recurrence of propagators.expm on its component of the
//     state; the other components come from the neighbouring lanes by shuffles (no shared
//     memory, no barrier in the recurrence);
//   * the generator is SHIFTED by the mid-range c0 of the drift's diagonal,
//     exp(f A h) = exp(c0 h) exp((f A - c0) h): the Taylor degree follows the norm of the
//     shifted matrix (bounded per step by its largest absolute row sum), which halves the
//     number of terms for ladder-like spectra (transmon); exp(c0 dt) comes from the pre-pass;
//   * records of the step after next are loaded while the current step computes.
// First order, one GPU; everything else stays with the kernels of kq_small.cuh / kq_warp.cuh.
#pragma once
#include "kq_small.cuh"
#include "kq_lanes_geom.cuh"

// objective, row and base lane of this lane
struct LnLane {
  int k, r, base;
  bool act;
};
__device__ __forceinline__ LnLane ln_lane(const KqLanes& d, int K, int N) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LnLane q;
  const int g = lane / d.NP;
  q.r = lane - g * d.NP;
  q.k = warp * d.G + g;
  q.base = lane - q.r;
  q.act = q.k < K && q.r < N;
  return q;
}

// Mid-range c0 of the diagonal of f * (sum of the drift terms) of objective k; identical in
// every lane of the objective (and in both kernels: same operations in the same order).
template <int FSEL>
__device__ __forceinline__ cplx ln_shift(const KqSweepArgs& a, const KqLanes& d, const LnLane& q) {
  const int N = a.N, M = a.M, NN = N * N;
  const int kk = min(q.k, a.K - 1), rr = min(q.r, N - 1);   // padding lanes shadow a real row
  cplx dg = c_zero();
  for (int m = 0; m < M; ++m) {   // (branch-free: the warp stays converged)
    const cplx v = apply_f<FSEL>(a.ops[((size_t)kk * M + m) * NN + rr * N + rr]);
    const bool drift = a.term2pulse[kk * M + m] == -1;
    dg = c_make(dg.x + (drift ? v.x : 0.0), dg.y + (drift ? v.y : 0.0));
  }
  double lo_x = dg.x, hi_x = dg.x, lo_y = dg.y, hi_y = dg.y;
  for (int off = d.NP >> 1; off > 0; off >>= 1) {
    lo_x = fmin(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, off));
    hi_x = fmax(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, off));
    lo_y = fmin(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, off));
    hi_y = fmax(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, off));
  }
  return c_make(0.5 * (lo_x + hi_x), 0.5 * (lo_y + hi_y));
}

// ---- pre-pass: grid NT, block W * 32 (the lane layout of the chain kernel) ---------------
template <int FSEL>
__global__ void __launch_bounds__(256) k_rows_prep(const KqSweepArgs a, const KqLanes d) {
  if (a.cond_epoch && *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch) return;
  const int K = a.K, N = a.N, NT = a.NT, L = a.L, NN = N * N;
  const int n = blockIdx.x, t = threadIdx.x, BT = blockDim.x;
  const LnLane q = ln_lane(d, K, N);
  const cplx c0 = ln_shift<FSEL>(a, d, q);
  const double dtn = a.dt[n];
  const double cn = q.act ? a.chi_norms[q.k] : 0.0;
  const cplx* chi = a.X + ((size_t)n * K + (q.act ? q.k : 0)) * N;
  cplx* zrow = d.zeta + (size_t)n * (L + 1) * BT;
  for (int l = 0; l < L; ++l) {
    cplx z = c_zero();
    if (q.act) {
      const cplx* mu = a.mu + ((size_t)q.k * L + l) * NN + (size_t)q.r * N;   // column r
      for (int r = 0; r < N; ++r) z = c_fma_conj(chi[r], mu[r], z);
    }
    zrow[(size_t)l * BT + t] = c_make(z.x * cn, z.y * cn);
  }
  {
    double s, c;
    sincos(c0.y * dtn, &s, &c);
    const double e = exp(c0.x * dtn);
    zrow[(size_t)L * BT + t] = c_make(e * c, e * s);   // exp(c0 dt)
  }
  if (t < KQ_LN_LMAX) {
    double* sc = d.scal + (size_t)n * KQ_LN_SC;
    const bool in = t < L;
    sc[t] = in ? a.shape[(size_t)t * NT + n] / a.lambda_a[t] : 0.0;   // optimize.py:474
    sc[KQ_LN_LMAX + t] = in ? a.pulses[(size_t)t * NT + n] : 0.0;
    if (t == 0) {
      sc[2 * KQ_LN_LMAX] = dtn;
      sc[2 * KQ_LN_LMAX + 1] = 0.0;
    }
  }
}

// ---- the chain ----------------------------------------------------------------------------
struct LnRec {
  cplx z[KQ_LN_LMAX];
  cplx ph;
  double2 s[KQ_LN_SC / 2];
};
__device__ __forceinline__ void ln_load(LnRec& R, const KqLanes& d, int n, int L, int BT) {
  const cplx* zrow = d.zeta + (size_t)n * (L + 1) * BT + threadIdx.x;
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) R.z[l] = (l < L) ? zrow[(size_t)l * BT] : c_zero();
  R.ph = zrow[(size_t)L * BT];
  const double2* sp = reinterpret_cast<const double2*>(d.scal + (size_t)n * KQ_LN_SC);
#pragma unroll
  for (int i = 0; i < KQ_LN_SC / 2; ++i) R.s[i] = sp[i];
}

template <int NZ>
struct LnState {
  cplx y;                          // this lane's component of phi_k
  double ga[KQ_LN_LMAX];
  cplx T[KQ_MMAX_SMALL][NZ];       // entries of row r of f * T_m (the first drift term: - c0)
  int src[NZ];                     // lane the entry's column lives in
  int t2p[KQ_MMAX_SMALL];
};

// Entries of row q.r of objective q.k: the non-zero columns (union over the terms, and the
// diagonal), f * T_m at those columns, the first drift term shifted by - c0 on the diagonal.
// Uniform control flow -- padding lanes shadow a real row and take nothing -- so that the warp
// provably stays converged for the shuffles of the time loop.  Returns true if the row has
// more than NZ entries.
template <int NZ, int FSEL>
__device__ __forceinline__ bool ln_setup(const KqSweepArgs& a, const LnLane& q, cplx c0,
                                         LnState<NZ>& S) {
  const int K = a.K, N = a.N, M = a.M, NN = N * N;
#pragma unroll
  for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
    S.t2p[m] = (m < M && q.act) ? a.term2pulse[q.k * M + m] : -2;
#pragma unroll
    for (int z = 0; z < NZ; ++z) S.T[m][z] = c_zero();
  }
#pragma unroll
  for (int z = 0; z < NZ; ++z) S.src[z] = threadIdx.x & 31;
  int cnt = 0;
  bool shifted = false, overflow = false;
  const int kk = min(q.k, K - 1), rr = min(q.r, N - 1);
  // the diagonal first (entry 0 is the lane's own component: no shuffle), then the others
  for (int cc = 0; cc < N; ++cc) {
    const int c = (cc == 0) ? rr : (cc <= rr ? cc - 1 : cc);
    cplx t[KQ_MMAX_SMALL];
    bool nz = (c == rr);
#pragma unroll
    for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
      t[m] = c_zero();
      if (m < M) {
        t[m] = a.ops[((size_t)kk * M + m) * NN + (size_t)c * N + rr];
        nz = nz || t[m].x != 0.0 || t[m].y != 0.0;
      }
    }
    nz = nz && q.act;
    const bool take = nz && cnt < NZ;
    overflow = overflow || (nz && cnt >= NZ);
#pragma unroll
    for (int z = 0; z < NZ; ++z) {
      const bool here = take && z == cnt;
      S.src[z] = here ? q.base + c : S.src[z];
#pragma unroll
      for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
        cplx v = apply_f<FSEL>(t[m]);
        const bool sh = here && c == rr && !shifted && S.t2p[m] == -1;
        if (sh) v = c_sub(v, c0);
        shifted = shifted || sh;
        S.T[m][z] = here ? v : S.T[m][z];
      }
    }
    cnt += take ? 1 : 0;
  }
  return overflow;
}

// Row of f A - c0 (times h) under the pulse values eps, Taylor plan from its largest absolute
// row sum in the warp, the Taylor/Horner recurrence on the lane's component, exp(c0 dt).
template <int NZ>
__device__ __forceinline__ cplx ln_propagate(const LnState<NZ>& S, cplx y0,
                                             const double (&eps)[KQ_LN_LMAX], double dtn,
                                             cplx phase, int M) {
  const KqTables& T = c_kq_tables;
  cplx A[NZ];
#pragma unroll
  for (int z = 0; z < NZ; ++z) A[z] = c_zero();
#pragma unroll
  for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
    if (m < M) {
      const int t = S.t2p[m];
      double coef = (t == -1) ? 1.0 : 0.0;
#pragma unroll
      for (int l = 0; l < KQ_LN_LMAX; ++l) coef = (t == l) ? eps[l] : coef;
#pragma unroll
      for (int z = 0; z < NZ; ++z) A[z] = c_fma_real(coef, S.T[m][z], A[z]);
    }
  }
  double x = 0.0;
#pragma unroll
  for (int z = 0; z < NZ; ++z) x += fabs(A[z].x) + fabs(A[z].y);
  x *= fabs(dtn);
  {
    // one Taylor plan per warp: the largest row sum of its objectives (rounded up)
    const int hi = __reduce_max_sync(0xffffffffu, __double2hiint(x));
    x = __hiloint2double(hi, (int)0xffffffff);
  }
  int s, mdeg;
  taylor_plan_fine(T, x, s, mdeg);
  const double h = (s == 1) ? dtn : dtn / (double)s;
#pragma unroll
  for (int z = 0; z < NZ; ++z) A[z] = c_make(h * A[z].x, h * A[z].y);
  cplx yc = y0;
  for (int rep = 0; rep < s; ++rep) {
    const cplx v = yc;
    cplx y = v;
    for (int j = mdeg; j >= 1; --j) {
      const double cj = T.inv[j];   // consumed at the end of the term: its latency is hidden
      cplx w0 = c_zero(), w1 = c_zero();
#pragma unroll
      for (int z = 0; z < NZ; ++z) {
        cplx yz = y;   // entry 0: the diagonal
        if (z > 0) {
          yz.x = __shfl_sync(0xffffffffu, y.x, S.src[z]);
          yz.y = __shfl_sync(0xffffffffu, y.y, S.src[z]);
        }
        if (z & 1) w1 = c_fma(A[z], yz, w1);
        else w0 = c_fma(A[z], yz, w0);
      }
      const cplx w = (NZ > 1) ? c_add(w0, w1) : w0;
      y = c_fma_real(cj, w, v);
    }
    yc = y;
  }
  return c_make(phase.x * yc.x - phase.y * yc.y, phase.x * yc.y + phase.y * yc.x);
}

template <int NZ, bool MULTI>
__device__ __forceinline__ void ln_step(const KqSweepArgs& a, const KqLanes& d, const LnRec& R,
                                        LnState<NZ>& S, double (*red)[KQ_LN_LMAX][8], int n,
                                        int L, int M) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double sl[KQ_LN_LMAX] = {R.s[0].x, R.s[0].y, R.s[1].x, R.s[1].y};
  const double g[KQ_LN_LMAX] = {R.s[2].x, R.s[2].y, R.s[3].x, R.s[3].y};
  const double dtn = R.s[4].x;
  // ---- Im <chi_k| mu_lk |phi_k> ||chi_k||, summed over rows and objectives
  double o[KQ_LN_LMAX];
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) o[l] = fma(R.z[l].x, S.y.y, R.z[l].y * S.y.x);
  // (straight-line butterfly over the whole warp: lanes without a row hold zeros, and branches
  // around shuffles cost more than the idle levels)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int l = 0; l < KQ_LN_LMAX; ++l) o[l] += __shfl_xor_sync(0xffffffffu, o[l], off);
  }
  if (MULTI) {   // objectives in several warps: one shared-memory hop (kept compact: the
                 // whole time step should stay in the instruction cache)
    const int par = n & 1;
    if (lane == 0) {
#pragma unroll
      for (int l = 0; l < KQ_LN_LMAX; ++l) red[par][l][warp] = o[l];
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < KQ_LN_LMAX; ++l) {
      double acc = red[par][l][0];
#pragma unroll 1
      for (int w = 1; w < d.W; ++w) acc += red[par][l][w];
      o[l] = acc;
    }
  }
  // ---- updated pulse values, rounded like optimize.py:474-477
  double eps[KQ_LN_LMAX];
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) {
    eps[l] = __dadd_rn(g[l], __dmul_rn(sl[l], o[l]));
    S.ga[l] = __dadd_rn(S.ga[l], __dmul_rn(__dmul_rn(sl[l], __dmul_rn(o[l], o[l])), dtn));
    if (threadIdx.x == 0 && l < L) a.opt_pulses[(size_t)l * a.NT + n] = eps[l];
  }
  // ---- forward step under the updated pulses
  S.y = ln_propagate<NZ>(S, S.y, eps, dtn, R.ph, M);
}

template <int NZ, int FSEL, bool MULTI>
__global__ void __launch_bounds__(256, 1) k_fwupd_rows(const KqSweepArgs a, const KqLanes d) {
  __shared__ double red[2][KQ_LN_LMAX][8];
  if (a.cond_epoch && *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch) return;
  const int K = a.K, N = a.N, NT = a.NT, M = a.M, L = a.L;
  const int BT = blockDim.x;
  const LnLane q = ln_lane(d, K, N);
  const cplx c0 = ln_shift<FSEL>(a, d, q);
  LnState<NZ> S;
  const bool overflow = ln_setup<NZ, FSEL>(a, q, c0, S);
  // kq_problem.row_nnz promised <= NZ entries per row: a violation is reported through the
  // status word (the host raises); no early exit, the warp stays whole
  if (overflow) atomicExch(a.status, (int)-5);
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) S.ga[l] = 0.0;
  S.y = q.act ? a.state0[(size_t)q.k * N + q.r] : c_zero();
  // records travel global -> registers two steps ahead of their use (one copy of the step
  // in the loop body: register moves are cheaper than instruction-cache misses here)
  LnRec Rc, Rn;
  ln_load(Rc, d, 0, L, BT);
  ln_load(Rn, d, min(1, NT - 1), L, BT);
#pragma unroll 1
  for (int n = 0; n < NT; ++n) {
    LnRec Rnn;
    ln_load(Rnn, d, min(n + 2, NT - 1), L, BT);
    ln_step<NZ, MULTI>(a, d, Rc, S, red, n, L, M);
    Rc = Rn;
    Rn = Rnn;
  }
  if (a.stateT && q.act) a.stateT[(size_t)q.k * N + q.r] = S.y;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int l = 0; l < KQ_LN_LMAX; ++l)
      if (l < L) a.g_a[l] = S.ga[l];
  }
}

// ---- propagation sweeps without update (optimize.py:806-886) for the same problems -------
// Tasks as in k_sweep_warp (kq_warp.cuh): one per objective, or -- time-parallel propagation,
// a.seg_pass 1 / 2 -- one per (objective, segment[, basis vector]); a task is a group of NP
// lanes, G tasks per warp.  The pulse values, dt and the phase of the shift are evaluated one
// step ahead of their use.
template <int NZ, int FSEL>
__global__ void __launch_bounds__(256, 1) k_prop_rows(const KqSweepArgs a, const KqLanes d) {
  const int K = a.K, N = a.N, NT = a.NT, M = a.M, L = a.L, NN = N * N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int seg_pass = a.seg_pass;
  const int n_task = seg_pass ? a.k_cnt * a.seg_count * (seg_pass == 1 ? N : 1) : a.k_cnt;
  const int g = lane / d.NP;
  int task = (blockIdx.x * nwarps + warp) * d.G + g;
  const bool valid = task < n_task;
  if (!valid) task = n_task - 1;
  int seg = 0, vec = 0;
  if (seg_pass) {
    const int rest = task / a.k_cnt;
    task -= rest * a.k_cnt;
    seg = rest % a.seg_count;
    vec = rest / a.seg_count;
  }
  LnLane q;
  q.r = lane - g * d.NP;
  q.k = a.k_lo + task;
  q.base = lane - q.r;
  q.act = valid && q.r < N;
  const cplx c0 = ln_shift<FSEL>(a, d, q);
  LnState<NZ> S;
  const bool overflow = ln_setup<NZ, FSEL>(a, q, c0, S);
  if (overflow && a.status) atomicExch(a.status, (int)-5);
  const int rr = min(q.r, N - 1);
  cplx y = c_zero();
  if (q.act) {
    if (seg_pass == 1) y = c_make(q.r == vec ? 1.0 : 0.0, 0.0);
    else if (seg_pass == 2) y = a.seg_B[((size_t)seg * K + q.k) * N + q.r];
    else y = a.state0[(size_t)q.k * N + q.r];
  }
  const bool bwd = a.backward != 0;
  int n_first = bwd ? NT - 1 : 0, n_count = NT;
  if (seg_pass) {   // time window of this segment, in sweep order
    if (bwd) {
      const int w1 = NT - seg * a.seg_len, w0 = max(0, w1 - a.seg_len);
      n_first = w1 - 1;
      n_count = w1 - w0;
    } else {
      const int w0 = seg * a.seg_len, w1 = min(NT, w0 + a.seg_len);
      n_first = w0;
      n_count = w1 - w0;
    }
  }
  const int n_step = bwd ? -1 : 1;
  if (a.store && q.act && seg == 0 && seg_pass != 1) {
    const size_t r0 = bwd ? (size_t)NT : 0;
    kq_store(a, (r0 * K + q.k) * N + q.r, y);
  }
  // scalars of the first step
  double eps[KQ_LN_LMAX];
#pragma unroll
  for (int l = 0; l < KQ_LN_LMAX; ++l) eps[l] = (l < L) ? a.pulses[(size_t)l * NT + n_first] : 0.0;
  double dtn = a.dt[n_first];
  (void)NN;
  (void)rr;
  for (int it = 0, n = n_first; it < n_count; ++it, n += n_step) {
    // next step's scalars (independent of the chain)
    const int nn = (it + 1 < n_count) ? n + n_step : n;
    double eps_next[KQ_LN_LMAX];
#pragma unroll
    for (int l = 0; l < KQ_LN_LMAX; ++l) eps_next[l] = (l < L) ? a.pulses[(size_t)l * NT + nn] : 0.0;
    const double dt_next = a.dt[nn];
    // exp(c0 dt) of the shift
    double sn, cs;
    sincos(c0.y * dtn, &sn, &cs);
    const double ex = (c0.x != 0.0) ? exp(c0.x * dtn) : 1.0;
    y = ln_propagate<NZ>(S, y, eps, dtn, c_make(ex * cs, ex * sn), M);
    if (a.store && q.act && seg_pass != 1) {
      const size_t r1 = bwd ? (size_t)n : (size_t)n + 1;
      kq_store(a, (r1 * K + q.k) * N + q.r, y);
    }
#pragma unroll
    for (int l = 0; l < KQ_LN_LMAX; ++l) eps[l] = eps_next[l];
    dtn = dt_next;
  }
  if (seg_pass == 1) {
    // column `vec` of the segment propagator (column-major)
    if (q.act) a.seg_P[((size_t)seg * K + q.k) * NN + (size_t)vec * N + q.r] = y;
  } else if (a.stateT && q.act && (seg_pass == 0 || seg == a.seg_count - 1)) {
    a.stateT[(size_t)q.k * N + q.r] = y;
  }
}
