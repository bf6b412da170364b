// Krotov iteration (optimize.py:393-508 of the reference) for FEW objectives with one
// control: the sequential update/forward chain is kept, but every link is ONE small
// matrix-vector product instead of a Taylor series of them, and everything else --
// including the backward sweep -- is parallel in time.
//
// The step propagator under the pulse eps_n = a_n + Delta_n (a: an ANCHOR pulse, a guess
// pulse of this or an earlier Krotov iteration) is entire in Delta,
//     U_n(Delta) = exp(f (T0 + (a_n + Delta) T1) dt_n) = sum_j Delta^j E_j[n],
// and |Delta| dt ||T1|| is small, so a handful of coefficient matrices E_0..E_J reproduce it
// to rounding (remainder <= x^(J+1)/(J+1)!, x = |Delta| dt ||T1||).  The E_j[n] depend on
// the anchor only.  They are computed for all time steps in parallel (k_dp_build: the
// Taylor/Horner recurrence of propagators.expm carried out on matrix polynomials in Delta,
// truncated at degree J) and KEPT in the caller's workspace: successive Krotov iterations
// change the pulse by less and less, so the records serve until the pulse has drifted out
// of the radius they were built for (k_dp_plan decides on the device; steady state: no
// build at all).  One Krotov iteration is then
//   plan     chi_k(T) (built-in constructors), norms, bounds, reuse / rebuild decision
//   [build]  E_j[n] around the current guess                          (parallel in n)
//   segprod  U_n = sum_j D_n^j E_j[n], D = guess - anchor, multiplied over segments of the
//            time grid                                                (parallel in segments)
//   expand   backward states chi[n] = U_n^dag chi[n+1] inside every segment, from the
//            segment's boundary state (= product of the later segment propagators applied
//            to chi(T));  zeta_kj[n+1] = E_j[n]^dag mu^dag chi_k[n+1] ||chi_k||; per-step
//            scalars                                                  (parallel in segments)
//   sweep    the sequential chain, one CTA: per time step
//              P = Horner_j(Delta_n; E_j[n])       element-wise, in registers
//              phi_{n+1} = P phi_n                 lanes = (objective, row, column group)
//              delta_{n+1} = (S/lambda) Im sum_j Delta_n^j sum_k <zeta_kj[n+1] | phi_k[n]>
//            -- the overlap with the backward state is evaluated from phi_n, one step
//            ahead, as one more row of the same matrix-vector product.  With <= 32 lanes the
//            chain runs in ONE warp on shuffles only (no barrier, no shared-memory hop).
//            One contiguous record per time step, laid out lane-major, is streamed
//            HBM -> shared memory by TMA bulk copies into a ring (producer warp) and from
//            there into registers one step ahead of its use.
//   epilogue tau_k, bookkeeping for the next call's plan.
// The sweep verifies |Delta| <= radius; a violation (or a problem the plan declines) is
// reported through status[1] / status[3] / diag_out like a fixed point that did not
// converge, and the caller repeats the iteration with the sequential Taylor kernels.
#pragma once
#include "kq_spec.cuh"

#include "kq_dpoly_geom.cuh"

// Record layout (complex numbers): lane-major with the degree fastest,
//   rec[(cc * NL + lane) * JS + j],  JS = (J + 1) | 1  (odd stride: conflict-free LDS.128),
// lane = (k * (N + 1) + r) * Q + q for row r of the AUGMENTED matrix polynomial of objective
// k -- rows 0..N-1 are E_j (written by the build kernel, kept across calls), row N is
// conj(zeta_j)^T (written by the expand kernel in every call) -- and column q * C + cc; then
// two numbers {S/lambda, guess} and {dt, guess - anchor} of this time step.
__device__ __forceinline__ int dp_js(int J) { return (J + 1) | 1; }
__device__ __forceinline__ int dp_rec_used(const KqDpoly& d, int J) {
  return d.C * d.NL * dp_js(J) + 2;
}
__device__ __forceinline__ size_t dp_rec_index(const KqDpoly& d, int NR, int k, int row, int col,
                                               int JS) {
  const int q = col / d.C, cc = col - q * d.C;
  return ((size_t)cc * d.NL + (size_t)(k * NR + row) * d.Q + q) * JS;
}
__device__ __forceinline__ bool dp_declined(const KqSweepArgs& a) {
  return *reinterpret_cast<volatile int*>(a.status + 1) == (int)a.epoch;
}

// ---- plan (one CTA): chi(T), bounds, reuse / rebuild ------------------------------
__global__ void __launch_bounds__(256) k_dp_plan(const KqSweepArgs a, const KqDpoly d) {
  __shared__ double red[8][8];
  __shared__ double sred[8];
  __shared__ double tsum_s[2];
  const int tid = threadIdx.x, K = a.K, N = a.N, NT = a.NT, NN = N * N;
  KqDpHeader* h = d.hdr;
  // chi_k(T) of functionals.py:177-197 (ss), 225-253 (sm), 293-317 (re), 389-437 (hs), then
  // chi / ||chi|| and ||chi|| (optimize.py:407-410)
  if (a.chi_kind >= 0) {
    if (a.chi_kind == KQ_CHI_SM) {
      double x = 0.0, y = 0.0;
      if (a.tau_sum) {
        x = a.tau_sum[0].x;
        y = a.tau_sum[0].y;
      } else {
        for (int j = tid; j < K; j += 256) {
          const double w = a.weights ? a.weights[j] : 1.0;
          x = fma(w, a.tau_in[j].x, x);
          y = fma(w, a.tau_in[j].y, y);
        }
        x = warp_allreduce_sum(x);
        y = warp_allreduce_sum(y);
        if ((tid & 31) == 0) {
          red[0][tid >> 5] = x;
          red[1][tid >> 5] = y;
        }
        __syncthreads();
        x = red[0][0];
        y = red[1][0];
        for (int w = 1; w < 8; ++w) {
          x += red[0][w];
          y += red[1][w];
        }
        __syncthreads();
      }
      if (tid == 0) {
        tsum_s[0] = x;
        tsum_s[1] = y;
      }
      __syncthreads();
    }
    const double Kt = (double)a.K_total;
    for (int k = tid; k < K; k += 256) {
      const double w = a.weights ? a.weights[k] : 1.0;
      cplx c;
      if (a.chi_kind == KQ_CHI_RE || a.chi_kind == KQ_CHI_HS) {
        c = c_make(w * (1.0 / (2.0 * Kt)), 0.0);
      } else if (a.chi_kind == KQ_CHI_SS) {
        const cplx t = a.tau_in[k];
        c = c_make(t.x / Kt * w, t.y / Kt * w);
      } else {
        const double f = (1.0 / (Kt * Kt)) * w;
        c = c_make(f * tsum_s[0], f * tsum_s[1]);
      }
      double nrm2 = 0.0;
      for (int i = 0; i < N; ++i) {
        cplx t = a.targets[(size_t)k * N + i];
        if (a.chi_kind == KQ_CHI_HS) t = c_sub(t, a.phiT_in[(size_t)k * N + i]);
        const cplx v = c_make(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
        d.chi[(size_t)k * N + i] = v;
        nrm2 = fma(v.x, v.x, nrm2);
        nrm2 = fma(v.y, v.y, nrm2);
      }
      const double nrm = sqrt(nrm2);
      d.norms[k] = nrm;
      for (int i = 0; i < N; ++i) {   // a target reached exactly gives chi = 0, not 0/0
        const cplx v = d.chi[(size_t)k * N + i];
        d.chi[(size_t)k * N + i] = nrm > 0.0 ? c_make(v.x / nrm, v.y / nrm) : c_zero();
      }
    }
    __syncthreads();
  }
  double dtmax = 0.0, gmax = 0.0, slmax = 0.0, o0 = 0.0, o1 = 0.0, Dmax = 0.0;
  const double lam = a.lambda_a[0];
  const bool have_anchor = h->anchor_epoch != 0;
  const bool have_last =
      h->valid_epoch != 0 && (uint32_t)h->valid_epoch + 1u == a.epoch && h->last_max >= 0.0;
  // dt and the operator norms belong to the problem: kept in the header after the first call
  const bool have_const = h->dtmax > 0.0;
  bool finite = true;
  // four entries per thread in flight (the loop is bound by the latency of its loads)
  for (int n0 = tid; n0 < NT; n0 += 4 * 256) {
    double g[4], an[4], dtv[4], sh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = n0 + u * 256;
      const bool in = n < NT;
      g[u] = in ? a.pulses[n] : 0.0;
      an[u] = (in && have_anchor) ? d.anchor[n] : g[u];
      dtv[u] = (in && !have_const) ? a.dt[n] : 0.0;
      sh[u] = (in && !have_last) ? a.shape[n] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      dtmax = fmax(dtmax, fabs(dtv[u]));
      gmax = fmax(gmax, fabs(g[u]));
      slmax = fmax(slmax, fabs(sh[u] / lam));
      Dmax = fmax(Dmax, fabs(g[u] - an[u]));
      if (!(fabs(g[u]) < 1e300)) finite = false;
    }
  }
  // a-priori bound of |Im sum_k <chi_k| mu |phi_k>| ||chi_k||: sum_k ||chi_k|| ||mu_k||_1 ||phi_k(0)||
  // (unitary or contractive dynamics); only needed while no update has been measured
  double ap = 0.0;
  if (!have_const || !have_last) {
    for (int k = tid; k < K; k += 256) {
      const int t2p = a.term2pulse[k * 2 + 1];
      const double n0 = a.op_norm[k * 2 + 0], n1 = a.op_norm[k * 2 + 1];
      if (t2p == 0) {
        o0 = fmax(o0, n0);
        o1 = fmax(o1, n1);
      } else if (t2p == -1) {
        o0 = fmax(o0, n0 + n1);
      } else {
        o0 = fmax(o0, n0);
      }
      if (!have_last) {
        double mun = 0.0;
        for (int c = 0; c < N; ++c) {
          double cs = 0.0;
          for (int r = 0; r < N; ++r) {
            const cplx m = a.mu[(size_t)k * NN + c * N + r];
            cs += fabs(m.x) + fabs(m.y);
          }
          mun = fmax(mun, cs);
        }
        double pn = 0.0;
        for (int r = 0; r < N; ++r) {
          const cplx s = a.state0[(size_t)k * N + r];
          pn += fabs(s.x) + fabs(s.y);
        }
        ap += d.norms[k] * mun * fmax(pn, 1.0);
      }
    }
  }
  double v[6] = {dtmax, gmax, slmax, o0, o1, Dmax};
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    v[j] = warp_allreduce_max(v[j]);
    if ((tid & 31) == 0) red[j][tid >> 5] = v[j];
  }
  ap = warp_allreduce_sum(ap);
  if ((tid & 31) == 0) sred[tid >> 5] = ap;
  const bool all_finite = __syncthreads_and(finite) != 0;
  if (tid != 0) return;
  for (int j = 0; j < 6; ++j)
    for (int w = 1; w < 8; ++w) red[j][0] = fmax(red[j][0], red[j][w]);
  for (int w = 1; w < 8; ++w) sred[0] += sred[w];
  gmax = red[1][0];
  slmax = red[2][0];
  Dmax = red[5][0];
  if (have_const) {
    dtmax = h->dtmax;
    o0 = h->o0;
    o1 = h->o1;
  } else {
    dtmax = red[0][0];
    o0 = red[3][0];
    o1 = red[4][0];
    h->dtmax = dtmax;
    h->o0 = o0;
    h->o1 = o1;
  }
  // predicted size of this iteration's update
  // (a strict a-priori bound before the first measured update)
  double upd = have_last ? 2.5 * h->last_max : slmax * sred[0];
  upd = fmax(upd, 1e-9 * fmax(gmax, 1e-3));
  bool ok = all_finite && (upd < 1e300);
  // smallest degree whose radius covers `want`: x^(J+1)/(J+1)! <= 2e-17.  With a measured
  // update size the records are built for GROW times the prediction, so that they serve
  // several iterations; the a-priori bound (first call) is loose enough as it is.  If even
  // the highest degree falls short it is used nevertheless: the sweep verifies the radius.
  const double x1 = dtmax * o1;   // x = |Delta| x1
  const double want = have_last ? KQ_DP_GROW * upd : upd;
  int J = 1;
  double radius = 1e300;
  if (x1 > 0.0) {
    // (2e-17 (J+1)!)^(1/(J+1)) for J = 1..8
    const double xj[KQ_DP_JMAX] = {6.324555320336759e-09, 4.9324241486609435e-06,
                                   1.4801656089845705e-04, 1.1913578981670911e-03,
                                   4.9324241486609416e-03, 1.3910780714804482e-02,
                                   3.078355801572517e-02,  5.78509288716415e-02};
    for (int jj = 1; jj <= KQ_DP_JMAX; ++jj) {
      J = jj;
      radius = xj[jj - 1] / x1;
      if (radius >= want) break;
    }
  }
  // reuse the records while the pulse stays inside their radius -- unless they carry much
  // higher a degree than this iteration needs (every degree costs time in the sweep)
  bool rebuild = true;
  if (ok && have_anchor && Dmax + upd <= h->radius && h->J <= J + 1) rebuild = false;
  if (ok && rebuild) {
    // the build kernel has no squaring stage: the whole step must be a plain Taylor series
    if (ok && !(dtmax * (o0 + (gmax + radius) * o1) <= 1.0)) {
      const double rmax = (x1 > 0.0) ? (1.0 / dtmax - o0) / o1 - gmax : -1.0;
      if (rmax >= 2.0 * upd && dtmax * (o0 + gmax * o1) <= 1.0)
        radius = fmin(radius, rmax);
      else
        ok = false;
    }
    if (ok) {
      int ps, pm;
      double pb;
      plan_bound(dtmax * (o0 + (gmax + fmin(radius, 1e290)) * o1), ps, pm, pb);
      h->J = J;
      h->m = pm;
      h->radius = radius;
      h->anchor_max = gmax;
      h->anchor_epoch = (int)a.epoch;
      h->builds += 1;
    }
  } else if (ok) {
    h->reuses += 1;
  }
  h->rebuild = (ok && rebuild) ? 1 : 0;
  h->usable = ok ? 1 : 0;
  if (!ok) {
    h->anchor_epoch = 0;
    a.status[1] = (int)a.epoch;   // the caller repeats with the sequential kernels
  }
}

// One Horner stage of the matrix polynomial for element (r, c): new_j = [j == 0] 1 +
// hi (A old_j + B old_{j-1}), DOLD = number of old degrees (compile time: no predicated work).
// px = old polynomial + c (element (x, c) of degree j at px[j * NN + x * N]), pn = new + r N + c.
template <int NMAX, int DOLD>
__device__ __forceinline__ void dp_stage(const cplx (&A)[NMAX], const cplx (&B)[NMAX],
                                         const cplx* px, cplx* pn, int N, int NN, int jtop,
                                         double hi, bool diag) {
  cplx y[DOLD + 1];
#pragma unroll
  for (int j = 0; j <= DOLD; ++j) y[j] = c_zero();
#pragma unroll
  for (int x = 0; x < NMAX; ++x) {
    if (x < N) {
#pragma unroll
      for (int j = 0; j < DOLD; ++j) {
        const cplx p = px[j * NN + x * N];
        y[j] = c_fma(A[x], p, y[j]);
        y[j + 1] = c_fma(B[x], p, y[j + 1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j <= DOLD; ++j)
    if (j <= jtop)
      pn[j * NN] = c_make(fma(hi, y[j].x, (j == 0 && diag) ? 1.0 : 0.0), hi * y[j].y);
}

// ---- build: the E part of the records of all time steps ---------------------------
// grid (ceil(NT / TPC), K), block TPC * N * N threads: thread = (time step, row r, column c)
// keeps row r of A = f (T0 + a_n T1) and of B = f T1 in registers and computes element
// (r, c) of every coefficient matrix; the matrix polynomial lives in shared memory
// ([2][j][x][c] per time step, double-buffered over the Horner stages).  All time steps use
// the Taylor degree m the plan kernel chose for the largest step.
template <int NMAX>
__global__ void __launch_bounds__(256) k_dp_build(const KqSweepArgs a, const KqDpoly d) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  if (!d.hdr->rebuild) return;
  const int N = a.N, NN = N * N, NT = a.NT, NR = N + 1;
  const int k = blockIdx.y, tid = threadIdx.x;
  const int ti = tid / NN, e = tid - ti * NN;
  const int r = e / N, c = e - r * N;
  const int n_raw = blockIdx.x * d.TPC + ti;
  const bool live = n_raw < NT + 2 * KQ_DP_PAD;
  const int n = live ? n_raw : NT - 1;   // idle slots shadow the last step (no writes)
  const bool pad = n >= NT;              // records behind the grid: identity steps (dt = 0)
  const int J = d.hdr->J, m = d.hdr->m;
  const int PS = (J + 1) * NN;                       // one polynomial
  cplx* P0 = reinterpret_cast<cplx*>(dp_smem) + (size_t)ti * (2 * PS);
  cplx* P1 = P0 + PS;
  const int t2p = a.term2pulse[k * 2 + 1];
  const bool driven = (t2p == 0);
  const double g = pad ? 0.0 : (driven ? a.pulses[n] : (t2p == -1 ? 1.0 : 0.0));
  const double h = pad ? 0.0 : a.dt[n];
  cplx A[NMAX], B[NMAX];
#pragma unroll
  for (int x = 0; x < NMAX; ++x) {
    A[x] = c_zero();
    B[x] = c_zero();
    if (x < N) {
      cplx t0 = a.ops[((size_t)k * 2 + 0) * NN + x * N + r];
      cplx t1 = a.ops[((size_t)k * 2 + 1) * NN + x * N + r];
      if (!a.is_super) {
        t0 = apply_f<0>(t0);
        t1 = apply_f<0>(t1);
      }
      A[x] = c_make(fma(g, t1.x, t0.x), fma(g, t1.y, t0.y));
      if (driven) B[x] = t1;
    }
  }
  for (int j = 0; j <= J; ++j) {
    P0[j * NN + r * N + c] = c_make((j == 0 && r == c) ? 1.0 : 0.0, 0.0);
    P1[j * NN + r * N + c] = c_zero();
  }
  __syncthreads();
  cplx* Pold = P0;
  cplx* Pnew = P1;
  for (int i = m; i >= 1; --i) {
    const double hi = h * c_kq_tables.inv[i];
    const int dold = min(J, m - i) + 1;   // degrees 0..dold-1 are present before this stage
    const int jtop = min(J, m - i + 1);   // degrees present after it
    const cplx* px = Pold + c;
    cplx* pn = Pnew + r * N + c;
    const bool diag = (r == c);
    switch (dold) {
      case 1: dp_stage<NMAX, 1>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      case 2: dp_stage<NMAX, 2>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      case 3: dp_stage<NMAX, 3>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      case 4: dp_stage<NMAX, 4>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      case 5: dp_stage<NMAX, 5>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      case 6: dp_stage<NMAX, 6>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      case 7: dp_stage<NMAX, 7>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      case 8: dp_stage<NMAX, 8>(A, B, px, pn, N, NN, jtop, hi, diag); break;
      default: dp_stage<NMAX, 9>(A, B, px, pn, N, NN, jtop, hi, diag); break;
    }
    __syncthreads();
    cplx* t = Pold;
    Pold = Pnew;
    Pnew = t;
  }
  // ---- record n: rows 0..N-1 = E_j, zeros in the padding columns
  if (!live) return;
  const int JS = dp_js(J), Npad = d.Npad;
  cplx* rec = d.rec + (size_t)n * dp_rec_used(d, J);
  const int per_obj = N * Npad * (J + 1);   // elements of this objective: (row, column, j)
  for (int o = e; o < per_obj; o += NN) {
    const int j = o % (J + 1);
    const int rc = o / (J + 1);
    const int col = rc % Npad, row = rc / Npad;
    const cplx v = (col < N) ? Pold[j * NN + row * N + col] : c_zero();
    rec[dp_rec_index(d, NR, k, row, col, JS) + j] = v;
  }
  if (e == 0 && k == 0 && !pad) d.anchor[n] = a.pulses[n];
}

// Stage the E coefficients of `count` records (n_first, n_first + dir, ...) of objective k into
// shared memory, dst[(s (J+1) + j) NN + r N + c], with cp.async (16 bytes each; every thread of
// the CTA takes part).  The caller commits / waits.
__device__ __forceinline__ void dp_stage_records(const KqDpoly& d, cplx* dst, int k, int N, int J,
                                                 int used, int n_first, int dir, int count) {
  const int NN = N * N, NR = N + 1, JS = dp_js(J), per = (J + 1) * NN;
  for (int idx = threadIdx.x; idx < count * per; idx += blockDim.x) {
    const int sidx = idx / per, rem = idx - sidx * per;
    const int j = rem / NN, el = rem - j * NN;
    const int r = el / N, c = el - r * N;
    const cplx* src =
        d.rec + (size_t)(n_first + dir * sidx) * used + dp_rec_index(d, NR, k, r, c, JS) + j;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + idx)), "l"(src)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory");
}

// ---- segprod: propagators of the time segments under the guess pulse ---------------
// grid (nseg, K), block >= N*N threads: thread = element (r, c).  P <- U_n P over the steps
// of the segment (ascending), U_n = sum_j D_n^j E_j[n]; output row-major.  The records come
// through shared memory in double-buffered batches of d.estage_cap / ((J+1) N N) steps.
__global__ void __launch_bounds__(256) k_dp_segprod(const KqSweepArgs a, const KqDpoly d) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  if (dp_declined(a)) return;
  const int N = a.N, NN = N * N, NT = a.NT, K = a.K;
  const int k = blockIdx.y, q = blockIdx.x, e = threadIdx.x;
  const bool act = e < NN;
  const int r = act ? e / N : 0, c = act ? e - r * N : 0;
  const int J = d.hdr->J, used = dp_rec_used(d, J);
  cplx* U = reinterpret_cast<cplx*>(dp_smem);   // [NN]
  cplx* P = U + NN;                              // [2][NN]
  cplx* Eb = P + 2 * NN;                         // [2][estage_cap]
  double* Ds = reinterpret_cast<double*>(Eb + (size_t)2 * d.estage_cap);   // [seg_len]
  const int n0 = q * d.seg_len, n1 = min(NT, n0 + d.seg_len);
  const int per = (J + 1) * NN;
  const int SB = max(1, d.estage_cap / per);
  const int nb = (n1 - n0 + SB - 1) / SB;
  const bool driven = a.term2pulse[k * 2 + 1] == 0;
  for (int i = e; i < n1 - n0; i += blockDim.x)
    Ds[i] = driven ? a.pulses[n0 + i] - d.anchor[n0 + i] : 0.0;
  if (act) P[e] = c_make(r == c ? 1.0 : 0.0, 0.0);
  dp_stage_records(d, Eb, k, N, J, used, n0, 1, min(SB, n1 - n0));
  cp_async_commit();
  int cur = 0;
  for (int b = 0; b < nb; ++b) {
    const int nf = n0 + b * SB, cnt = min(SB, n1 - nf);
    if (b + 1 < nb) {
      dp_stage_records(d, Eb + (size_t)((b + 1) & 1) * d.estage_cap, k, N, J, used, nf + SB, 1,
                       min(SB, n1 - nf - SB));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const cplx* Ebb = Eb + (size_t)(b & 1) * d.estage_cap;
    for (int si = 0; si < cnt; ++si) {
      const double D = Ds[nf + si - n0];
      cplx u = c_zero();
      if (act) {
        const cplx* ep = Ebb + (size_t)si * per + e;
        u = ep[J * NN];
        for (int j = J - 1; j >= 0; --j) {
          const cplx ej = ep[j * NN];
          u.x = fma(u.x, D, ej.x);
          u.y = fma(u.y, D, ej.y);
        }
        U[e] = u;
      }
      __syncthreads();
      if (act) {
        cplx a0 = c_zero(), a1 = c_zero();
        const cplx* Pc = P + cur * NN + c;
        const cplx* Ur = U + r * N;
        int x = 0;
        for (; x + 1 < N; x += 2) {
          a0 = c_fma(Ur[x], Pc[x * N], a0);
          a1 = c_fma(Ur[x + 1], Pc[(x + 1) * N], a1);
        }
        if (x < N) a0 = c_fma(Ur[x], Pc[x * N], a0);
        P[(cur ^ 1) * NN + e] = c_add(a0, a1);
      }
      __syncthreads();
      cur ^= 1;
    }
  }
  if (act) d.segP[((size_t)q * K + k) * NN + e] = P[cur * NN + e];
}

// ---- expand: backward states, zeta rows and per-step scalars ------------------------
// grid (nseg, K), block >= N * R2 threads: thread = (column c, row r), r fastest (R2 = N
// rounded up to a power of two, so that a column's rows are neighbouring lanes of one warp).
// Boundary state of the segment: chi(T) pulled back through the later segments,
// v <- P_p^dag v for p = nseg-1 .. q+1 (thread r < N forms component r from column r of P_p).
// Then, for n = n1-1 .. n0,
//   eta = mu^dag chi[n+1] ||chi||,  zeta_j[c] = sum_r E_j[n][r, c] conj(eta_r)   -> record n
//   chi[n][c] = sum_r conj(U_n[r, c]) chi[n+1][r]                                -> X[n]
// (sums over r: butterfly over the R2 lanes of the column).  With d.chain == 0 the backward
// states are read from X instead (kq_sweep_forward_update: the caller propagated them).
// Segment propagators and records come through shared memory in double-buffered batches.
template <int NMAX>
__global__ void __launch_bounds__(256) k_dp_expand(const KqSweepArgs a, const KqDpoly d) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  if (dp_declined(a)) return;
  const int N = a.N, NN = N * N, NT = a.NT, NR = N + 1, K = a.K, R2 = d.R2;
  const int k = blockIdx.y, q = blockIdx.x, tid = threadIdx.x;
  const int c = tid / R2, r = tid - c * R2;
  const bool act = (c < N) && (r < N);
  const int J = d.hdr->J, JS = dp_js(J), used = dp_rec_used(d, J);
  cplx* schi = reinterpret_cast<cplx*>(dp_smem);   // [2][N]
  cplx* seta = schi + 2 * N;                        // [N] conj(eta)
  cplx* Pb = seta + N;                              // [2][chain_bs][NN]
  cplx* Eb = Pb + (size_t)2 * d.chain_bs * NN;      // [2][estage_cap]
  cplx* mus = Eb + (size_t)2 * d.estage_cap;        // [NN] mu of this objective
  double* Ds = reinterpret_cast<double*>(mus + NN); // [seg_len] guess - anchor
  const int n0 = q * d.seg_len, n1 = min(NT, n0 + d.seg_len);
  const int per = (J + 1) * NN;
  const int SB = max(1, d.estage_cap / per);
  const int nbe = (n1 - n0 + SB - 1) / SB;
  const double cnorm = d.norms[k];
  const bool driven = a.term2pulse[k * 2 + 1] == 0;
  const double lam = a.lambda_a[0];
  for (int i = tid; i < NN; i += blockDim.x) mus[i] = a.mu[(size_t)k * NN + i];
  for (int i = tid; i < n1 - n0; i += blockDim.x) {
    const int n = n0 + i;
    const double D = driven ? a.pulses[n] - d.anchor[n] : 0.0;
    Ds[i] = D;
    if (k == 0) {   // this step's scalars
      cplx* tail = d.rec + (size_t)n * used + (size_t)d.C * d.NL * JS;
      tail[0] = c_make(a.shape[n] / lam, a.pulses[n]);
      tail[1] = c_make(a.dt[n], D);
    }
  }
  // the first batch of records is under way while the boundary state is formed
  dp_stage_records(d, Eb, k, N, J, used, n1 - 1, -1, min(SB, n1 - n0));
  cp_async_commit();
  int cur = 0;
  if (d.chain) {
    if (tid < N) schi[tid] = d.chi[(size_t)k * N + tid];
    __syncthreads();
    const int BS = d.chain_bs, nchain = d.nseg - 1 - q;
    const int nb = (nchain + BS - 1) / BS;
    const bool row_thread = tid < N;
    auto issue = [&](int b) {
      const int cnt = min(BS, nchain - b * BS);
      cplx* dst = Pb + (size_t)(b & 1) * BS * NN;
      for (int e = tid; e < cnt * NN; e += blockDim.x) {
        const int i = e / NN, el = e - i * NN;
        const int pseg = d.nseg - 1 - (b * BS + i);
        const cplx* src = d.segP + ((size_t)pseg * K + k) * NN + el;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + i * NN + el)),
                     "l"(src)
                     : "memory");
      }
      cp_async_commit();
    };
    if (nb > 0) issue(0);
    for (int b = 0; b < nb; ++b) {
      if (b + 1 < nb) {
        issue(b + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const int cnt = min(BS, nchain - b * BS);
      const cplx* Pbb = Pb + (size_t)(b & 1) * BS * NN;
      for (int i = 0; i < cnt; ++i) {
        if (row_thread) {
          const cplx* Pp = Pbb + i * NN + tid;   // column `tid` of the row-major propagator
          cplx a0 = c_zero(), a1 = c_zero();
          int x = 0;
          for (; x + 1 < N; x += 2) {
            a0 = c_fma_conj(Pp[x * N], schi[cur * N + x], a0);
            a1 = c_fma_conj(Pp[(x + 1) * N], schi[cur * N + x + 1], a1);
          }
          if (x < N) a0 = c_fma_conj(Pp[x * N], schi[cur * N + x], a0);
          schi[(cur ^ 1) * N + tid] = c_add(a0, a1);
        }
        __syncthreads();
        cur ^= 1;
      }
    }
    if (q == d.nseg - 1 && tid < N) d.X[((size_t)NT * K + k) * N + tid] = schi[cur * N + tid];
  }
  const int el = (act ? r : 0) * N + (act ? c : 0);
  for (int b = 0; b < nbe; ++b) {
    const int nf = n1 - 1 - b * SB, cnt = min(SB, nf - n0 + 1);
    if (b + 1 < nbe) {
      dp_stage_records(d, Eb + (size_t)((b + 1) & 1) * d.estage_cap, k, N, J, used, nf - SB, -1,
                       min(SB, nf - SB - n0 + 1));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const cplx* Ebb = Eb + (size_t)(b & 1) * d.estage_cap;
    for (int si = 0; si < cnt; ++si) {
      const int n = nf - si;
      if (!d.chain) {
        if (tid < N) schi[cur * N + tid] = d.X[((size_t)(n + 1) * K + k) * N + tid];
        __syncthreads();
      }
      if (tid < N) {
        // conj(eta), eta = mu^dag chi[n+1] ||chi||; no update follows the last step
        const double cn = (n + 1 < NT) ? cnorm : 0.0;
        cplx acc = c_zero();
        for (int rr = 0; rr < N; ++rr)
          acc = c_fma_conj(mus[tid * N + rr], schi[cur * N + rr], acc);
        seta[tid] = c_make(acc.x * cn, -acc.y * cn);
      }
      __syncthreads();
      const double D = Ds[n - n0];
      cplx E[KQ_DP_JMAX + 1];
      const cplx* ep = Ebb + (size_t)si * per + el;
#pragma unroll
      for (int j = 0; j <= KQ_DP_JMAX; ++j) E[j] = (act && j <= J) ? ep[j * NN] : c_zero();
      cplx u = E[KQ_DP_JMAX];
#pragma unroll
      for (int j = KQ_DP_JMAX - 1; j >= 0; --j) {
        u.x = fma(u.x, D, E[j].x);
        u.y = fma(u.y, D, E[j].y);
      }
      const cplx chr = act ? schi[cur * N + r] : c_zero();
      const cplx etr = act ? seta[r] : c_zero();
      cplx px = c_fma_conj(u, chr, c_zero());   // conj(U[r,c]) chi[r]
      cplx pz[KQ_DP_JMAX + 1];
#pragma unroll
      for (int j = 0; j <= KQ_DP_JMAX; ++j) pz[j] = c_fma(E[j], etr, c_zero());
      for (int o = R2 >> 1; o > 0; o >>= 1) {
        px.x += __shfl_xor_sync(0xffffffffu, px.x, o);
        px.y += __shfl_xor_sync(0xffffffffu, px.y, o);
#pragma unroll
        for (int j = 0; j <= KQ_DP_JMAX; ++j) {
          if (j <= J) {
            pz[j].x += __shfl_xor_sync(0xffffffffu, pz[j].x, o);
            pz[j].y += __shfl_xor_sync(0xffffffffu, pz[j].y, o);
          }
        }
      }
      cplx* rec = d.rec + (size_t)n * used;
      if (r == 0 && c < N) {
        if (d.chain) {
          schi[(cur ^ 1) * N + c] = px;
          d.X[((size_t)n * K + k) * N + c] = px;
        }
        cplx* zp = rec + dp_rec_index(d, NR, k, N, c, JS);
#pragma unroll
        for (int j = 0; j <= KQ_DP_JMAX; ++j)
          if (j <= J) zp[j] = pz[j];
        if (c < d.Npad - N) {   // padding columns of the zeta row
          cplx* z0 = rec + dp_rec_index(d, NR, k, N, N + c, JS);
          for (int j = 0; j <= J; ++j) z0[j] = c_zero();
        }
      }
      __syncthreads();
      if (d.chain) cur ^= 1;
    }
  }
  if (q == d.nseg - 1) {
    // records behind the grid: no overlap, no update
    for (int i = tid; i < 2 * KQ_DP_PAD * d.Npad; i += blockDim.x) {
      const int n = NT + i / d.Npad, col = i % d.Npad;
      cplx* zp = d.rec + (size_t)n * used + dp_rec_index(d, NR, k, N, col, JS);
      for (int j = 0; j <= J; ++j) zp[j] = c_zero();
    }
    if (k == 0)
      for (int i = tid; i < 2 * KQ_DP_PAD; i += blockDim.x) {
        cplx* tail = d.rec + (size_t)(NT + i) * used + (size_t)d.C * d.NL * JS;
        tail[0] = c_zero();
        tail[1] = c_zero();
      }
  }
}

// ---- the sequential sweep (one CTA) -------------------------------------------------
// Threads 0..NLP-1 (NLP = lanes rounded up to whole warps) are the consumers; one more warp
// is the producer: its lane 0 refills the ring (waits until every consumer warp has released
// a stage, re-arms the stage's mbarrier and issues the TMA bulk copy), so no copy is issued
// from the dependency chain.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dp_consumer_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// The chain is issue-bound as much as latency-bound (ONE warp, or a few in lock step): what
// counts is the number of instructions per time step.  Hence
//  * the ring holds CHUNKS of S consecutive records (one bulk copy, one mbarrier round trip
//    per chunk); inside a chunk the next record is a constant offset away;
//  * records beyond the time grid are identity steps, so the loop needs no tail handling
//    (the step count is rounded up to a multiple of S; only the pulse store is predicated);
//  * the row sums are reduce-scattered over the Q lanes of a row (first butterfly level splits
//    Re / Im between the half groups: half the shuffles of an all-reduce) and spread through
//    shared memory with ONE store per lane (lanes that own nothing write to a dummy slot: no
//    divergence) and a warp-level (or named-barrier) sync; the two exchange buffers alternate
//    with the two halves of the loop body, so their addresses are loop constants;
//  * |Delta| <= radius is checked on the high words (integer max).
struct DpSweepCtx {
  uint64_t* full;
  uint64_t* empty;
  double* xb;          // [2][XS] exchange buffers: phi (2 K Npad) | zeta sums (2 K) | dummy slots
  const cplx* ring;
  int K, NT, NL, NLP, R, S, used, XS;
  int lane_id;
  int wr;              // double index (in a buffer) of this lane's store
  int rd_phi;          // complex index of this lane's first column of phi_k
  int d_off, d_str;    // zeta sums: double index of objective 0's Im, stride
};

template <int C, int J>
struct DpCoef {
  cplx e[C][J + 1];
};
struct DpScal {
  double sl, g, dt, D;
};

template <int C, int J>
__device__ __forceinline__ void dp_load(DpCoef<C, J>& R, const cplx* lane_rec, int colblk) {
#pragma unroll
  for (int cc = 0; cc < C; ++cc)
#pragma unroll
    for (int j = 0; j <= J; ++j) R.e[cc][j] = lane_rec[cc * colblk + j];
}
__device__ __forceinline__ void dp_load_scal(DpScal& S, const cplx* tail) {
  const cplx t0 = tail[0], t1 = tail[1];
  S.sl = t0.x;
  S.g = t0.y;
  S.dt = t1.x;
  S.D = t1.y;
}

struct DpChain {
  double Delta;   // a_n + Delta = pulse of the step about to be taken
  double ga;
  int mx;         // max over the steps of the high word of |Delta|
  int left;       // pulse values still to be written
  double* optp;   // where the next one goes
};

// One time step: `Rn` = coefficients of record n, `Sx` = scalars of record n+1 (they give the
// next update).  ph[cc] = this lane's columns of phi_n on entry, of phi_{n+1} on exit.
// xw = the exchange buffer of this half of the loop body.  PFN: the coefficients of the
// record after next are loaded into `Rp` from `pp`, one load behind every Horner level --
// while the FP64 pipe works on the chain the shared-memory pipe is idle; a batch of loads
// in front of the shuffles would delay them by its length.
template <int C, int J, bool QONE, bool ONEWARP, bool PFN>
__device__ __forceinline__ void dp_step(const DpSweepCtx& c, const DpCoef<C, J>& Rn,
                                        DpCoef<C, J>& Rp, const cplx* pp, int colblk,
                                        const DpScal& Sx, cplx (&ph)[C], DpChain& s, double* xw,
                                        int Q, int qbit) {
  cplx acc = c_zero();
#pragma unroll
  for (int cc = 0; cc < C; ++cc) {
    cplx Pc = Rn.e[cc][J];
    if (PFN) Rp.e[cc][J] = pp[cc * colblk + J];
#pragma unroll
    for (int j = J - 1; j >= 0; --j) {
      Pc.x = fma(Pc.x, s.Delta, Rn.e[cc][j].x);
      Pc.y = fma(Pc.y, s.Delta, Rn.e[cc][j].y);
      if (PFN) Rp.e[cc][j] = pp[cc * colblk + j];
    }
    acc = c_fma(Pc, ph[cc], acc);
  }
  if (QONE) {
    *reinterpret_cast<cplx*>(xw + c.wr) = acc;
  } else {
    double v = qbit ? acc.y : acc.x;
    const double snd = qbit ? acc.x : acc.y;
    v += __shfl_xor_sync(0xffffffffu, snd, Q >> 1);
    if (Q > 2) {
      if (Q > 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
    }
    xw[c.wr] = v;
  }
  if (ONEWARP)
    __syncwarp();
  else
    dp_consumer_sync(c.NLP);
  // sum of the zeta rows over the objectives (the slots of objectives >= K hold zeros)
  const double* db = xw + c.d_off;
  double dsum;
  if (c.K <= 4) {
    const double d0 = db[0], d1 = db[c.d_str], d2 = db[2 * c.d_str], d3 = db[3 * c.d_str];
    dsum = (d0 + d1) + (d2 + d3);
  } else {
    dsum = db[0];
    for (int kk = 1; kk < c.K; ++kk) dsum += db[kk * c.d_str];
  }
  const cplx* phi = reinterpret_cast<const cplx*>(xw) + c.rd_phi;
#pragma unroll
  for (int cc = 0; cc < C; ++cc) ph[cc] = phi[cc];
  // delta_{n+1} = (S/lambda) Im sum_k <eta_k[n+1] | phi_k[n+1]>, already evaluated from phi_k[n]
  const double delta = Sx.sl * dsum;
  s.Delta = Sx.D + delta;
  s.ga = fma(Sx.sl * (dsum * dsum), Sx.dt, s.ga);
  s.mx = max(s.mx, __double2hiint(s.Delta) & 0x7fffffff);
  --s.left;
  if (threadIdx.x == 0 && s.left >= 0) *s.optp = Sx.g + delta;
  ++s.optp;
}

// The consumers' time loop for a compile-time degree J and C columns per lane.  PF: records
// go ring -> registers one step ahead (two register sets, loop body = two steps); otherwise
// (large CTAs, few registers) the coefficients are read from the ring in the step that uses
// them.
template <int C, int J, bool QONE, bool ONEWARP, bool PF>
__device__ __forceinline__ void dp_run(const KqSweepArgs& a, const DpSweepCtx& c, cplx (&ph)[C],
                                       DpChain& s, double sl0, double d0, int Q, int qbit) {
  constexpr int JS = (J + 1) | 1;
  const int R = c.R, S = c.S, used = c.used;
  const int colblk = c.NL * JS;
  const int stage = S * used;                        // complex numbers per ring stage
  const cplx* lane0 = c.ring + (size_t)c.lane_id * JS;
  const int tail_off = used - 2 - c.lane_id * JS;    // from a lane pointer to the record's scalars
  const bool releaser = (threadIdx.x & 31) == 0;
  double* xw1 = c.xb + c.XS;   // first step of a pair writes buffer 1, the second buffer 0
  double* xw0 = c.xb;
  DpCoef<C, J> RA, RB;
  DpScal SA, SB;
  mbar_wait(&c.full[0], 0);
  dp_load<C, J>(RA, lane0, colblk);
  dp_load_scal(SA, lane0 + tail_off);
  // first update: delta_0 = (S_0/lambda) d0 from chi[0] and phi(0) directly
  {
    const double delta = sl0 * d0;
    s.Delta = SA.D + delta;
    s.ga = sl0 * (d0 * d0) * SA.dt;
    s.mx = __double2hiint(s.Delta) & 0x7fffffff;
    s.left = c.NT - 1;
    s.optp = a.opt_pulses + 1;
    if (threadIdx.x == 0) a.opt_pulses[0] = SA.g + delta;
  }
  const int NC = (c.NT + S - 1) / S;   // chunks; chunk NC (identity records) is only looked at
  int st = 0;
  uint32_t phase = 0;
  // INPL (one warp): a step refills its own coefficient registers, level by level as its
  // Horner recurrence has consumed them, with the record two steps ahead.  Several warps
  // (shared-memory bandwidth matters more than the chain): the next record is loaded as a
  // batch in front of the step.
  constexpr bool INPL = PF && ONEWARP;
  constexpr bool BATCH = PF && !ONEWARP;
  if (INPL) dp_load<C, J>(RB, lane0 + used, colblk);
  for (int ch = 0; ch < NC; ++ch) {
    const cplx* p = lane0 + (size_t)st * stage;   // record of the step about to be taken
    int stn = st + 1;
    uint32_t phn = phase;
    if (stn == R) {
      stn = 0;
      phn ^= 1u;
    }
    const cplx* pn = lane0 + (size_t)stn * stage;
    for (int i = 0; i < S; i += 2) {
      // on entry RA = record n (at p) [INPL: and RB = record n+1]; p2 = record n+2: the
      // next ring stage for the last pair of a chunk
      const bool last = (i + 2 == S);
      const cplx* p2 = p + 2 * used;
      if (last) {
        mbar_wait(&c.full[stn], phn);
        p2 = pn;
      }
      if (!PF && (ch > 0 || i > 0)) dp_load<C, J>(RA, p, colblk);
      if (BATCH) dp_load<C, J>(RB, p + used, colblk);
      dp_load_scal(SB, p + used + tail_off);
      dp_step<C, J, QONE, ONEWARP, INPL>(c, RA, RA, p2, colblk, SB, ph, s, xw1, Q, qbit);
      if (!PF) dp_load<C, J>(RB, p + used, colblk);
      if (BATCH) dp_load<C, J>(RA, p2, colblk);
      dp_load_scal(SA, p2 + tail_off);
      dp_step<C, J, QONE, ONEWARP, INPL>(c, RB, RB, p2 + used, colblk, SA, ph, s, xw0, Q, qbit);
      p = p2;
    }
    __syncwarp();
    if (releaser) mbar_arrive(&c.empty[st]);
    st = stn;
    phase = phn;
  }
}

// MAXT: largest CTA of the instantiation (consumer lanes + the producer warp); PF: register
// prefetch of the records (needs the registers of a small CTA); QONE: one lane per row
template <int C, bool QONE, bool ONEWARP, int MAXT, bool PF>
__global__ void __launch_bounds__(MAXT, 1) k_dp_sweep(const KqSweepArgs a, const KqDpoly d) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  if (dp_declined(a)) return;
  const int N = a.N, NN = N * N, NT = a.NT, K = a.K, NR = N + 1;
  const int tid = threadIdx.x, BT = blockDim.x;
  const int NL = d.NL, Q = d.Q, Npad = d.Npad;
  const int NLP = BT - 32;        // consumer threads (whole warps)
  const int J = d.hdr->J;
  const int used_ = dp_rec_used(d, J);
  // chunks of S records (even), at least two ring stages, three if they fit
  int S = (d.ring / (3 * used_)) & ~1;
  S = max(2, min(KQ_DP_PAD, S));
  const int R = min(KQ_DP_RINGMAX, d.ring / (S * used_));
  DpSweepCtx c;
  c.lane_id = min(tid, NL - 1);   // consumer threads beyond NL shadow the last lane
  const bool act = tid < NL;
  const int lk = c.lane_id / (NR * Q), lr = (c.lane_id / Q) % NR, lq = c.lane_id % Q;
  c.K = K;
  c.NT = NT;
  c.NL = NL;
  c.NLP = NLP;
  c.R = R;
  c.S = S;
  c.used = used_;
  c.XS = 2 * K * Npad + 2 * K + 2 * NLP;
  c.full = reinterpret_cast<uint64_t*>(dp_smem);                              // [RINGMAX]
  c.empty = c.full + KQ_DP_RINGMAX;                                           // [RINGMAX]
  c.xb = reinterpret_cast<double*>(c.empty + KQ_DP_RINGMAX);                  // [2][XS]
  double* sred = c.xb + 2 * c.XS;                                             // [K*N]
  cplx* ring = reinterpret_cast<cplx*>(sred + ((K * N + 1) & ~1));            // [R][S][used]
  c.ring = ring;
  const uint32_t chunk_bytes = (uint32_t)(S * used_) * 16u;
  const int NC = (NT + S - 1) / S;

  if (tid == 0) {
    for (int st = 0; st < R; ++st) {
      mbar_init(&c.full[st], 1);
      mbar_init(&c.empty[st], (uint32_t)(NLP >> 5));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * c.XS; i += BT) c.xb[i] = 0.0;
  __syncthreads();
  if (tid >= NLP) {
    // ---- producer warp: chunks 0..NC (the last one is only looked at) ----
    if (tid != NLP) return;
    for (int ch = 0; ch <= NC; ++ch) {
      const int st = ch % R, use = ch / R;
      if (use > 0) mbar_wait(&c.empty[st], (uint32_t)((use - 1) & 1));
      mbar_expect_tx(&c.full[st], chunk_bytes);
      bulk_g2s(ring + (size_t)st * S * used_, d.rec + (size_t)ch * S * used_, chunk_bytes,
               &c.full[st]);
    }
    return;
  }
  // ---- consumers ----
  // phi(0) and the first overlap, d0 = Im sum_k <eta_k[0] | phi_k(0)>
  cplx* phi0 = reinterpret_cast<cplx*>(c.xb);
  if (act && lq == 0 && lr < N) {
    const cplx p0 = a.state0[(size_t)lk * N + lr];
    phi0[lk * Npad + lr] = p0;
    cplx eta = c_zero();
    for (int rr = 0; rr < N; ++rr)
      eta = c_fma_conj(a.mu[(size_t)lk * NN + lr * N + rr], d.X[(size_t)lk * N + rr], eta);
    sred[lk * N + lr] = d.norms[lk] * c_im_conj_mul(eta, p0);
  }
  // where this lane's row sum goes: Re from the lane q = 0, Im from the lane q = Q/2 (Q = 1:
  // the complex sum from the one lane of the row); everybody else has a dummy slot
  const int hq = Q >> 1;
  const int dbase = 2 * K * Npad;
  c.wr = dbase + 2 * K + 2 * tid;
  if (act) {
    if (lr < N) {
      if (lq == 0) c.wr = 2 * (lk * Npad + lr);
      if (Q > 1 && lq == hq) c.wr = 2 * (lk * Npad + lr) + 1;
    } else if (lq == hq) {
      c.wr = QONE ? dbase + 2 * lk : dbase + lk;
    }
  }
  c.d_off = QONE ? dbase + 1 : dbase;
  c.d_str = QONE ? 2 : 1;
  c.rd_phi = lk * Npad + lq * C;
  dp_consumer_sync(NLP);
  double d0 = 0.0;
  for (int i = 0; i < K * N; ++i) d0 += sred[i];
  const double sl0 = a.shape[0] / a.lambda_a[0];
  cplx ph[C];
#pragma unroll
  for (int cc = 0; cc < C; ++cc) ph[cc] = phi0[c.rd_phi + cc];
  const int qbit = (lq & hq) ? 1 : 0;
  DpChain s;
  switch (J) {
    case 1: dp_run<C, 1, QONE, ONEWARP, PF && (C * 2 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
    case 2: dp_run<C, 2, QONE, ONEWARP, PF && (C * 3 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
    case 3: dp_run<C, 3, QONE, ONEWARP, PF && (C * 4 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
    case 4: dp_run<C, 4, QONE, ONEWARP, PF && (C * 5 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
    case 5: dp_run<C, 5, QONE, ONEWARP, PF && (C * 6 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
    case 6: dp_run<C, 6, QONE, ONEWARP, PF && (C * 7 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
    case 7: dp_run<C, 7, QONE, ONEWARP, PF && (C * 8 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
    default: dp_run<C, 8, QONE, ONEWARP, PF && (C * 9 <= 20)>(a, c, ph, s, sl0, d0, Q, qbit); break;
  }
  // |Delta| <= radius for all steps, compared on the high words (conservative; NaN fails)
  const bool good = s.mx < (__double2hiint(d.hdr->radius) & 0x7fffffff);
  // the step count is even: the last step wrote exchange buffer 0
  if (good && act && lq == 0 && lr < N && a.stateT)
    a.stateT[(size_t)lk * N + lr] = phi0[lk * Npad + lr];
  if (tid == 0) {
    if (good) {
      a.g_a[0] = s.ga;
      a.status[2] = 0;
    } else {
      a.status[1] = (int)a.epoch;   // the series was built for smaller updates
    }
  }
}

// ---- epilogue: tau, status words for the caller, the largest update of this call
__global__ void __launch_bounds__(256) k_dp_epilogue(const KqSweepArgs a, const KqDpoly d,
                                                     int fallback_in_stream) {
  const bool failed = dp_declined(a);
  if (threadIdx.x == 0) {
    if (failed && !fallback_in_stream) atomicCAS(a.status + 3, 0, (int)a.epoch);
    if (a.diag_out) {
      a.diag_out[0] = *reinterpret_cast<volatile int*>(a.status);
      a.diag_out[1] = (failed && !fallback_in_stream) ? (int)a.epoch : 0;
      a.diag_out[2] = 0;
      a.diag_out[3] = 0;
    }
  }
  if (failed && !fallback_in_stream) return;   // outputs are not valid: the caller repeats
  if (a.tau_out && a.targets && a.stateT) {
    for (int k = threadIdx.x; k < a.K; k += 256) {
      cplx acc = c_zero();
      for (int i = 0; i < a.N; ++i)
        acc = c_fma_conj(a.targets[(size_t)k * a.N + i], a.stateT[(size_t)k * a.N + i], acc);
      a.tau_out[k] = acc;
    }
  }
  // the largest update of this call: predicts the next one (k_dp_plan)
  __shared__ double red[8];
  double m = 0.0;
  for (int n0 = threadIdx.x; n0 < a.NT; n0 += 4 * 256) {
    double o[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = n0 + u * 256;
      o[u] = n < a.NT ? a.opt_pulses[n] : 0.0;
      g[u] = n < a.NT ? a.pulses[n] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) m = fmax(m, fabs(o[u] - g[u]));
  }
  m = warp_allreduce_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
    d.hdr->last_max = m;
    d.hdr->valid_epoch = (int)a.epoch;
  }
}
