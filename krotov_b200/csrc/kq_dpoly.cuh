// Update/forward sweep (optimize.py:449-500 of the reference) for FEW objectives
// with one control: the sequential chain is kept, but every link is ONE small
// matrix-vector product instead of a Taylor series of them.
//
// The step propagator under the updated pulse eps_n = g_n + delta_n (g: guess
// pulse, delta: this iteration's update) is entire in delta,
//     U_n(delta) = exp(f (T0 + (g_n + delta) T1) dt_n) = sum_j delta^j E_j[n],
// and |delta| dt ||T1|| is tiny, so a handful of coefficient matrices E_0..E_J
// reproduce it to rounding (remainder <= x^(J+1)/(J+1)!, x = |delta| dt ||T1||).
// The E_j[n] depend on the guess pulse only and are computed for all time steps
// in parallel (k_dpoly_build: the Taylor/Horner recurrence of propagators.expm
// carried out on matrix polynomials in delta, truncated at degree J).  The
// sequential sweep (k_dpoly_sweep, one CTA) then does, per time step,
//     P = Horner_j(delta_n; E_j[n])        element-wise, in registers
//     phi_{n+1} = P phi_n                  lanes = (objective, row, column group)
//     delta_{n+1} = (S/lambda) Im sum_j delta_n^j sum_k <zeta_kj[n+1] | phi_k[n]>
// where zeta_kj[n+1] = E_j[n]^dag mu^dag chi_k[n+1] ||chi_k|| (k_dpoly_zeta, parallel
// over the time steps) -- the overlap with the backward state is evaluated from
// phi_n, one step ahead, so that it is not on the dependency chain.  One
// contiguous record per time step (E | zeta | next step's S/lambda, guess, dt),
// laid out lane-major, is streamed HBM -> shared memory by TMA bulk copies into
// a ring.  The degree J and a bound on |delta| are chosen on the device
// (k_dpoly_plan: 2.5 times the largest update of the Krotov iteration before,
// or an a-priori bound); the sweep verifies the bound and otherwise asks for the
// sequential Taylor kernels queued behind it (status[1] = epoch).
#pragma once
#include "kq_spec.cuh"

#include "kq_dpoly_geom.cuh"

// Record layout (complex numbers): lane-major with the degree fastest,
//   rec[(cc * NL + lane) * JS + j],  JS = (J + 1) | 1  (odd stride: conflict-free LDS.128),
// lane = (k * (N + 1) + r) * Q + q for row r of the AUGMENTED matrix polynomial of objective
// k -- rows 0..N-1 are E_j, row N is conj(zeta_j)^T -- and column q * C + cc; then two
// numbers {S/lambda, guess} and {dt, 0} of the NEXT time step.
__device__ __forceinline__ int dp_js(int J) { return (J + 1) | 1; }
__device__ __forceinline__ int dp_rec_used(const KqDpoly& d, int J) {
  return d.C * d.NL * dp_js(J) + 2;
}

// ---- plan: degree J and delta bound (one CTA) -------------------------------
__global__ void __launch_bounds__(256) k_dpoly_plan(const KqSweepArgs a, const KqDpoly d) {
  __shared__ double red[5][8];
  const int tid = threadIdx.x, K = a.K, N = a.N, NT = a.NT, NN = N * N;
  double dtmax = 0.0, gmax = 0.0, slmax = 0.0, o0 = 0.0, o1 = 0.0;
  const double lam = a.lambda_a[0];
  for (int n = tid; n < NT; n += 256) {
    dtmax = fmax(dtmax, fabs(a.dt[n]));
    gmax = fmax(gmax, fabs(a.pulses[n]));
    slmax = fmax(slmax, fabs(a.shape[n] / lam));
  }
  // a-priori bound of |Im sum_k <chi_k| mu |phi_k>| ||chi_k||: sum_k ||chi_k|| ||mu_k||_1 ||phi_k(0)||
  // (unitary or contractive dynamics)
  double ap = 0.0;
  for (int k = tid; k < K; k += 256) {
    o0 = fmax(o0, a.op_norm[k * 2 + 0]);
    const int t2p = a.term2pulse[k * 2 + 1];
    if (t2p == 0) o1 = fmax(o1, a.op_norm[k * 2 + 1]);
    if (t2p == -1) o0 = fmax(o0, a.op_norm[k * 2 + 0] + a.op_norm[k * 2 + 1]);
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
    ap += a.chi_norms[k] * mun * fmax(pn, 1.0);
  }
  double v[5] = {dtmax, gmax, slmax, o0, o1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    v[j] = warp_allreduce_max(v[j]);
    if ((tid & 31) == 0) red[j][tid >> 5] = v[j];
  }
  ap = warp_allreduce_sum(ap);
  __shared__ double reds[8];
  if ((tid & 31) == 0) reds[tid >> 5] = ap;
  __syncthreads();
  if (tid != 0) return;
  for (int j = 0; j < 5; ++j)
    for (int w = 1; w < 8; ++w) red[j][0] = fmax(red[j][0], red[j][w]);
  for (int w = 1; w < 8; ++w) reds[0] += reds[w];
  dtmax = red[0][0];
  gmax = red[1][0];
  slmax = red[2][0];
  o0 = red[3][0];
  o1 = red[4][0];
  double bound = slmax * reds[0];
  KqDpHeader* h = d.hdr;
  if (h->valid_epoch != 0 && (uint32_t)h->valid_epoch + 1u == a.epoch && h->last_max >= 0.0)
    bound = fmin(bound, 2.5 * h->last_max);
  bound = fmax(bound, 1e-9 * fmax(gmax, 1e-3));
  // remainder of the delta series relative to 1, per step: x^(J+1)/(J+1)! <= 2e-17
  const double x = bound * dtmax * o1;
  int J = 0;
  if (x > 0.0) {
    double term = x;   // x^(J+1)/(J+1)! for J = 0
    J = KQ_DP_JMAX + 1;
    for (int j = 0; j <= KQ_DP_JMAX; ++j) {
      if (term <= 2e-17) {
        J = j;
        break;
      }
      term *= x / (double)(j + 2);
    }
    if (J < 1) J = 1;
  }
  bool ok = J <= KQ_DP_JMAX;
  // the build kernel has no squaring stage: the whole step must be a plain Taylor series
  if (!(dtmax * (o0 + (gmax + bound) * o1) <= 1.0)) ok = false;
  if (!(bound < 1e300)) ok = false;
  int ps, pm;
  double pb;
  plan_bound(dtmax * (o0 + (gmax + bound) * o1), ps, pm, pb);
  h->J = ok ? J : 0;
  h->m = pm;
  h->delta_bound = bound;
  if (!ok) a.status[1] = (int)a.epoch;   // ask for the sequential kernels
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

// ---- build: the records of all time steps -----------------------------------
// grid (ceil(NT / TPC), K), block TPC * N * N threads: thread = (time step, row r, column c)
// keeps row r of A = f (T0 + g_n T1) and of B = f T1 in registers and computes element
// (r, c) of every coefficient matrix; the matrix polynomial lives in shared memory
// ([2][j][x][c] per time step, double-buffered over the Horner stages).  After the last
// stage the same threads form the zeta row from the backward state chi[n+1] and write the
// record.  All time steps use the Taylor degree m the plan kernel chose for the largest step.
template <int NMAX>
__global__ void __launch_bounds__(256) k_dpoly_build(const KqSweepArgs a, const KqDpoly d) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  if (*reinterpret_cast<volatile int*>(a.status + 1) == (int)a.epoch) return;
  const int N = a.N, NN = N * N, NT = a.NT, K = a.K, NR = N + 1;
  const int k = blockIdx.y, tid = threadIdx.x;
  const int ti = tid / NN, e = tid - ti * NN;
  const int r = e / N, c = e - r * N;
  const int n_raw = blockIdx.x * d.TPC + ti;
  const bool live = n_raw < NT;
  const int n = live ? n_raw : NT - 1;   // idle slots shadow the last step (no writes)
  const int J = d.hdr->J, m = d.hdr->m;
  const int PS = (J + 1) * NN;                       // one polynomial
  cplx* P0 = reinterpret_cast<cplx*>(dp_smem) + (size_t)ti * (2 * PS + N);
  cplx* P1 = P0 + PS;
  cplx* seta = P1 + PS;                              // conj(eta) of this step [N]
  const int t2p = a.term2pulse[k * 2 + 1];
  const bool driven = (t2p == 0);
  const double g = driven ? a.pulses[n] : (t2p == -1 ? 1.0 : 0.0);
  const double h = a.dt[n];
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
  if (e < N) {
    // conj(eta), eta = mu^dag chi[n+1] ||chi||; no update follows the last step
    const double cn = (n + 1 < NT) ? a.chi_norms[k] : 0.0;
    const cplx* chi = a.X + ((size_t)(n + 1) * K + k) * N;
    cplx acc = c_zero();
    for (int rr = 0; rr < N; ++rr) acc = c_fma_conj(a.mu[(size_t)k * NN + e * N + rr], chi[rr], acc);
    seta[e] = c_make(acc.x * cn, -acc.y * cn);
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
  // ---- record n: rows 0..N-1 = E_j, row N = conj(zeta_j)^T, zeros in the padding columns;
  // next step's scalars at the end
  if (!live) return;
  cplx* rec = d.rec + (size_t)n * d.rec_stride;
  const int Q = d.Q, C = d.C, NL = d.NL, JS = dp_js(J), Npad = d.Npad;
  const int per_obj = NR * Npad * (J + 1);   // elements of this objective: (row, column, j)
  for (int o = e; o < per_obj; o += NN) {
    const int j = o % (J + 1);
    const int rc = o / (J + 1);
    const int col = rc % Npad, row = rc / Npad;
    const int q = col / C, cc = col - q * C;
    cplx v = c_zero();
    if (col < N) {
      if (row < N) {
        v = Pold[j * NN + row * N + col];
      } else {
        for (int rr = 0; rr < N; ++rr) v = c_fma(Pold[j * NN + rr * N + col], seta[rr], v);
      }
    }
    rec[((size_t)cc * NL + (k * NR + row) * Q + q) * JS + j] = v;
  }
  if (e == 0 && k == 0) {
    cplx* tail = rec + (size_t)C * NL * JS;
    const bool nxt = n + 1 < NT;
    tail[0] = nxt ? c_make(a.shape[n + 1] / a.lambda_a[0], a.pulses[n + 1]) : c_zero();
    tail[1] = nxt ? c_make(a.dt[n + 1], 0.0) : c_zero();
  }
}

// ---- the sequential sweep (one CTA) -------------------------------------------
// Threads 0..NLP-1 (NLP = lanes rounded up to whole warps) are the consumers; one more warp
// is the producer: its lane 0 refills the ring (waits until every consumer warp has released
// a stage, re-arms the stage's mbarrier and issues the TMA bulk copy), so no copy is issued
// from the dependency chain.
// shared: [full 16][empty 16][dbuf 2 x K, padded][sred K N, padded][phi 2 x K Npad cplx][ring]
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dp_consumer_sync(int nthreads) {
  if (nthreads <= 32)
    __syncwarp();
  else
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

struct DpSweepCtx {
  uint64_t* full;
  uint64_t* empty;
  double* dbuf;
  cplx* sphi;
  const cplx* ring;
  int K, N, NT, NL, NLP, Q, Npad, R, stage, used;   // stage: complex numbers between ring stages
  int lane_id, k, r, q;
  bool act;
  uint32_t rec_bytes;
};

// The consumers' time loop for a compile-time degree J and C columns per lane.
template <int C, int J>
__device__ __forceinline__ void dp_run(const KqSweepArgs& a, const DpSweepCtx& c, double& delta,
                                       double& ga, double& dmax, int& cur) {
  constexpr int JS = (J + 1) | 1;
  const int K = c.K, N = c.N, NT = c.NT, R = c.R, Q = c.Q;
  const int PHS = K * c.Npad;                 // one phi buffer
  const int colblk = c.NL * JS;               // complex numbers between column blocks
  const cplx* lane_rec = c.ring + (size_t)c.lane_id * JS;
  const int phi_off = c.k * c.Npad + c.q * C;
  const bool is_phi_row = c.act && c.q == 0 && c.r < N;
  const bool is_d_row = c.act && c.q == 0 && c.r == N;
  const int phi_dst = c.k * c.Npad + c.r;
  const bool releaser = (threadIdx.x & 31) == 0;
  int st = 0;
  uint32_t phase = 0;
  for (int n = 0; n < NT; ++n) {
    mbar_wait(&c.full[st], phase);
    const cplx* e = lane_rec + (size_t)st * c.stage;
    const cplx* phi = c.sphi + cur * PHS + phi_off;
    cplx acc = c_zero();
#pragma unroll
    for (int cc = 0; cc < C; ++cc) {
      const cplx* ec = e + cc * colblk;
      cplx Pc = ec[J];
#pragma unroll
      for (int j = J - 1; j >= 0; --j) {
        const cplx ej = ec[j];
        Pc.x = fma(Pc.x, delta, ej.x);
        Pc.y = fma(Pc.y, delta, ej.y);
      }
      acc = c_fma(Pc, phi[cc], acc);
    }
    // next step's scalars travel with this record
    const cplx* tail = c.ring + (size_t)st * c.stage + (c.used - 2);
    const cplx t0 = tail[0], t1 = tail[1];
    for (int off = Q >> 1; off > 0; off >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    }
    // this warp is done with stage st
    __syncwarp();
    if (releaser) mbar_arrive(&c.empty[st]);
    const int nxt = cur ^ 1;
    if (is_phi_row) c.sphi[nxt * PHS + phi_dst] = acc;
    if (is_d_row) c.dbuf[nxt * K + c.k] = acc.y;
    dp_consumer_sync(c.NLP);
    if (n + 1 < NT) {
      // delta_{n+1} = (S/lambda) Im sum_k <eta_k[n+1] | phi_k[n+1]>, already evaluated from phi_k[n]
      const double* db = c.dbuf + nxt * K;
      double dsum = db[0];
      for (int kk = 1; kk < K; ++kk) dsum += db[kk];
      const double sl = t0.x;
      delta = sl * dsum;
      ga += sl * (dsum * dsum) * t1.x;
      dmax = fmax(dmax, fabs(delta));
      if (threadIdx.x == 0) a.opt_pulses[n + 1] = t0.y + delta;
    }
    cur = nxt;
    if (++st == R) {
      st = 0;
      phase ^= 1u;
    }
  }
}

template <int C>
__global__ void __launch_bounds__(KQ_DP_MAXLANES + 32, 1)
k_dpoly_sweep(const KqSweepArgs a, const KqDpoly d) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  if (*reinterpret_cast<volatile int*>(a.status + 1) == (int)a.epoch) return;
  const int N = a.N, NN = N * N, NT = a.NT, K = a.K, NR = N + 1;
  const int tid = threadIdx.x, BT = blockDim.x;
  const int NL = d.NL, Q = d.Q, Npad = d.Npad;
  const int NLP = BT - 32;        // consumer threads (whole warps)
  const int J = d.hdr->J;
  // as many ring stages as the records of this degree allow: the copies in flight hide the
  // latency of a bulk copy (about 1.5 us, i.e. several time steps)
  const int used_ = dp_rec_used(d, J);
  const int R = min(KQ_DP_RINGMAX, d.ring / used_);
  const double bound = d.hdr->delta_bound;
  DpSweepCtx c;
  c.lane_id = min(tid, NL - 1);   // consumer threads beyond NL shadow the last lane
  c.act = tid < NL;
  c.k = c.lane_id / (NR * Q);
  c.r = (c.lane_id / Q) % NR;
  c.q = c.lane_id % Q;
  c.K = K;
  c.N = N;
  c.NT = NT;
  c.NL = NL;
  c.NLP = NLP;
  c.Q = Q;
  c.Npad = Npad;
  c.R = R;
  c.stage = used_;
  c.full = reinterpret_cast<uint64_t*>(dp_smem);                              // [RINGMAX]
  c.empty = c.full + KQ_DP_RINGMAX;                                           // [RINGMAX]
  c.dbuf = reinterpret_cast<double*>(c.empty + KQ_DP_RINGMAX);                // [2][K]
  double* sred = c.dbuf + ((2 * K + 1) & ~1);                                 // [K*N]
  c.sphi = reinterpret_cast<cplx*>(sred + ((K * N + 1) & ~1));                // [2][K*Npad]
  cplx* ring = c.sphi + 2 * K * Npad;                                         // [R][rec_stride]
  c.ring = ring;
  c.used = used_;
  c.rec_bytes = (uint32_t)c.used * 16u;

  if (tid == 0) {
    for (int st = 0; st < R; ++st) {
      mbar_init(&c.full[st], 1);
      mbar_init(&c.empty[st], (uint32_t)(NLP >> 5));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * K * Npad; i += BT) c.sphi[i] = c_zero();
  __syncthreads();
  if (tid >= NLP) {
    // ---- producer warp ----
    if (tid != NLP) return;
    for (int n = 0; n < NT; ++n) {
      const int st = n % R, use = n / R;
      if (use > 0) mbar_wait(&c.empty[st], (uint32_t)((use - 1) & 1));
      mbar_expect_tx(&c.full[st], c.rec_bytes);
      bulk_g2s(ring + (size_t)st * used_, d.rec + (size_t)n * d.rec_stride, c.rec_bytes,
               &c.full[st]);
    }
    return;
  }
  // ---- consumers ----
  // phi(0) and the first update, delta_0 = (S_0/lambda) Im sum_k <eta_k[0] | phi_k(0)>
  if (c.act && c.q == 0 && c.r < N) {
    const int k = c.k, r = c.r;
    const cplx p0 = a.state0[(size_t)k * N + r];
    c.sphi[k * Npad + r] = p0;
    cplx eta = c_zero();
    for (int rr = 0; rr < N; ++rr)
      eta = c_fma_conj(a.mu[(size_t)k * NN + r * N + rr], a.X[(size_t)k * N + rr], eta);
    sred[k * N + r] = a.chi_norms[k] * c_im_conj_mul(eta, p0);
  }
  dp_consumer_sync(NLP);
  double d0 = 0.0;
  for (int i = 0; i < K * N; ++i) d0 += sred[i];
  const double sl0 = a.shape[0] / a.lambda_a[0];
  double delta = sl0 * d0;
  double ga = sl0 * (d0 * d0) * a.dt[0];
  double dmax = fabs(delta);
  if (tid == 0) a.opt_pulses[0] = a.pulses[0] + delta;
  int cur = 0;
  switch (J) {
    case 0: dp_run<C, 0>(a, c, delta, ga, dmax, cur); break;
    case 1: dp_run<C, 1>(a, c, delta, ga, dmax, cur); break;
    case 2: dp_run<C, 2>(a, c, delta, ga, dmax, cur); break;
    case 3: dp_run<C, 3>(a, c, delta, ga, dmax, cur); break;
    case 4: dp_run<C, 4>(a, c, delta, ga, dmax, cur); break;
    case 5: dp_run<C, 5>(a, c, delta, ga, dmax, cur); break;
    case 6: dp_run<C, 6>(a, c, delta, ga, dmax, cur); break;
    case 7: dp_run<C, 7>(a, c, delta, ga, dmax, cur); break;
    default: dp_run<C, 8>(a, c, delta, ga, dmax, cur); break;
  }
  if (c.act && c.q == 0 && c.r < N && a.stateT)
    a.stateT[(size_t)c.k * N + c.r] = c.sphi[cur * K * Npad + c.k * Npad + c.r];
  if (tid == 0) {
    KqDpHeader* h = d.hdr;
    if (dmax <= bound) {
      a.g_a[0] = ga;
      a.status[2] = 0;
      h->last_max = dmax;
      h->valid_epoch = (int)a.epoch;
    } else {
      a.status[1] = (int)a.epoch;   // the series was built for smaller updates: sequential kernels
    }
  }
}

// ---- after the conditional sequential kernels: keep the largest update of this call
// (the sweep kernel did that itself if it ran to the end); status words for the caller
__global__ void __launch_bounds__(256) k_dpoly_epilogue(const KqSweepArgs a, const KqDpoly d) {
  if (threadIdx.x == 0 && a.diag_out) {
    a.diag_out[0] = *reinterpret_cast<volatile int*>(a.status);
    a.diag_out[1] = 0;   // the sequential kernels are queued in-stream: nothing to repeat
    a.diag_out[2] = 0;
    a.diag_out[3] = 0;
  }
  if (*reinterpret_cast<volatile int*>(a.status + 1) != (int)a.epoch) return;
  __shared__ double red[8];
  double m = 0.0;
  for (int n = threadIdx.x; n < a.NT; n += 256) m = fmax(m, fabs(a.opt_pulses[n] - a.pulses[n]));
  m = warp_allreduce_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
    d.hdr->last_max = m;
    d.hdr->valid_epoch = (int)a.epoch;
  }
}
