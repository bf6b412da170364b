// Thread-per-objective kernels for small state vectors (N <= 4).
//
// One thread owns one objective: its state, the assembled generator A (N*N
// complex, registers) and the Horner/Taylor recurrence all stay in registers;
// the generator terms and mu live in shared memory laid out [term][elem][thread]
// (conflict-free 128-bit LDS).  The cross-objective sum of the pulse update
// (optimize.py:454-470 of the reference) is a warp butterfly + one shared
// memory hop + one __syncthreads per time step; with more objectives than one
// CTA holds, CTAs exchange their partial sums through flag-tagged 16-byte
// slots in global memory (kq_common.cuh) inside a co-resident grid.
#pragma once
#include "kq_common.cuh"

// largest CTA the fused kernel is compiled for: the BTMAX = 256 variant may
// use up to 255 registers per thread (no spills on the sequential chain), the
// 1024-thread variant (N <= 2 only) trades registers for objectives per CTA.
#define KQ_SMALL_MAXBT(N) ((N) <= 2 ? 1024 : 256)

struct KqSweepArgs {
  int K, N, NT, L, M, is_super;
  const cplx* ops;        // generator used by this sweep (ops or ops_adj)
  const cplx* mu;
  const int* term2pulse;
  const double* op_norm;
  const double* dt;
  const double* shape;
  const double* lambda_a;
  const double* pulses;   // guess pulses [L][NT]
  double* opt_pulses;     // [L][NT] (fw/update sweep)
  const cplx* state0;     // [K][N]
  cplx* stateT;           // [K][N] or null
  cplx* store;            // [NT+1][K][N] or null (prop: written, fwupd: Phi1)
  const cplx* X;          // [NT+1][K][N] backward states (fwupd)
  const double* chi_norms;
  const double* sigma;    // [NT] or null
  const cplx* Phi0;       // [NT+1][K][N] or null
  double* g_a;            // [L]
  int* status;            // workspace status word
  KqSlot* slots;          // cross-CTA exchange slots [2][gridDim.x][L]
  KqSlot* const* peer_slots;  // per-rank exchange buffers (device array) or null
  int rank, world;
  uint32_t tag_base;
  int backward;
  // propagation sweeps only: objectives [k_lo, k_lo + k_cnt) are propagated
  // (K stays the row stride of the stores)
  int k_lo, k_cnt;
  // time-parallel propagation (kq_spec.cuh): the sweep is cut into segments of
  // seg_len steps (sweep order); pass 1 computes the N x N propagator of every
  // segment into seg_P [S][K][N*N], a chain kernel turns them into the states
  // at the segment boundaries seg_B [S+1][K][N], pass 2 propagates every
  // segment from its boundary state and stores all states.  seg_pass 0 = off.
  int seg_len, seg_pass;
  int seg_count;      // number of segments (lane-per-row kernels: part of the task index)
  cplx* seg_P;
  cplx* seg_B;
  // time-parallel fused sweep (kq_picard.cuh): CTA = pic_Q objectives x pic_TC
  // time chunks of pic_W steps; cross-CTA exchange through tagged slots
  int pic_Q, pic_TC, pic_W, pic_maxit, pic_stride;
  double pic_rtol;
  KqSlot* pic_part;   // [owner][cta][2^pic_lwc] per-CTA partial sums over its objectives
  KqSlot* pic_eps;    // [cta][pic_stride]: mailbox of updated pulse values [owner][2^pic_lwc];
                      // CTA 0's mailbox continues with the owners' g_a shares [gridDim.x]
  // sequential kernels launched as the fall-back of the time-parallel sweep run
  // only if status[1] == cond_epoch (0 = unconditional)
  uint32_t cond_epoch, epoch;
  int pic_timing;     // 1: CTA 0 writes per-phase cycle counts to status[16..]
  int pic_window;     // index of the time window of a windowed update sweep
  int pic_accumulate; // 1: add g_a and the round count to the values of the window before
  // fused Krotov iteration (k_krotov_picard): chi boundary and backward sweep
  // inside the kernel
  int pic_lw, pic_Wc;        // log2(pic_W); time steps reduced per CTA (even)
  int pic_lwc;               // log2 of the slice length padded to a power of two >= 8
  int pic_bw;                // 1: backward sweep in the kernel (chi from chi_kind / chiT), 0: chi from X
  int chi_kind, K_total;     // KQ_CHI_* or -1 (normalised chiT + chi_norms given)
  const cplx* ops_adj;
  const cplx* chiT;
  const cplx* targets;
  const double* weights;
  const cplx* tau_in;
  cplx* tau_out;
  const cplx* phiT_in;
  const double* pic_hint;    // [NT] guess pulse of the iteration before, or null
  // history of the last Krotov updates in the caller's workspace (kq_krotov_iteration):
  // header {key = address the next guess is expected at, count, head} + ring of
  // 4 x pic_hist_ld doubles; the first iterate extrapolates the update from it
  unsigned long long* pic_hist_hdr;
  double* pic_hist;
  int pic_hist_ld;
  cplx* Xout;                // [NT+1][K][N] or null
  cplx* chi_out;             // [K][N] or null
  double* chi_norms_out;     // [K] or null
  int* diag_out;             // [4] copy of the status words {status, failed epoch, rounds, 0} or null
  // objectives sharded over GPUs (world > 1): after the owner CTA of a time slice has summed
  // the partial sums of this GPU's CTAs it writes that sum into the slice's slots of every
  // rank's exchange buffer (peer_slots[r] + pic_xg_off, layout [4][cta][rank][2^pic_lwc],
  // the 4 sub-buffers selected by the parities of epoch and round) and adds what the other
  // ranks wrote into its own buffer, in rank order: identical pulses on every GPU
  size_t pic_xg_off;         // in slots
  size_t pic_xg_buf;         // slots per sub-buffer
  const cplx* tau_sum;       // chis_sm with sharded objectives: sum_j w_j tau_j over ALL ranks
};

__device__ __forceinline__ void kq_store(const KqSweepArgs& a, size_t idx, cplx v) {
  a.store[idx] = v;
}

// y <- exp(f * A * dt) y,   A column-major N x N in registers.
template <int N, int FSEL>
__device__ __forceinline__ void expmv_small(const KqTables& T, const cplx (&A)[N * N],
                                            cplx (&y)[N], double dt, double x) {
  int s, m;
  double xs;
  taylor_plan(T, x, s, m, xs);
  const double h = (s == 1) ? dt : dt / (double)s;
  for (int rep = 0; rep < s; ++rep) {
    cplx v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = y[i];
    for (int j = m; j >= 1; --j) {
      const double cj = h * T.inv[j];
      cplx w[N];
#pragma unroll
      for (int r = 0; r < N; ++r) {
        cplx acc = c_zero();
#pragma unroll
        for (int c = 0; c < N; ++c) acc = c_fma(A[c * N + r], y[c], acc);
        w[r] = acc;
      }
#pragma unroll
      for (int r = 0; r < N; ++r) y[r] = c_fma_real(cj, apply_f<FSEL>(w[r]), v[r]);
    }
  }
}

// Shared-memory carve-up shared by both small kernels.
template <int N>
struct SmallSmem {
  cplx* ops;     // [M][N*N][BT]
  cplx* mu;      // [L][N*N][BT]
  double* red;   // [2][KQ_LMAX][32]
  double* tot;   // [2][KQ_LMAX]
  __device__ SmallSmem(unsigned char* raw, int M, int L, int BT, bool with_mu) {
    ops = reinterpret_cast<cplx*>(raw);
    mu = ops + (size_t)M * N * N * BT;
    double* d = reinterpret_cast<double*>(mu + (with_mu ? (size_t)L * N * N * BT : 0));
    red = d;
    tot = d + 2 * KQ_LMAX * 32;
  }
  static size_t bytes(int M, int L, int BT, bool with_mu) {
    return ((size_t)M + (with_mu ? L : 0)) * N * N * BT * sizeof(cplx) +
           (2 * KQ_LMAX * 32 + 2 * KQ_LMAX) * sizeof(double);
  }
};

// ---------------------------------------------------------------------------
// Propagation sweep without update: initial forward propagation
// (optimize.py:806-846) and backward propagation (optimize.py:849-886).
template <int N, int FSEL>
__global__ void __launch_bounds__(256, 1)
k_prop_small(const KqSweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const KqTables& T = c_kq_tables;
  constexpr int NN = N * N;
  const int BT = blockDim.x, tid = threadIdx.x;
  const int k = a.k_lo + blockIdx.x * BT + tid;
  const int K = a.K, NT = a.NT, M = a.M;
  if (k >= a.k_lo + a.k_cnt) return;  // no block-level synchronisation in this kernel
  SmallSmem<N> sm(smem_raw, M, a.L, BT, false);

  int t2p[KQ_MMAX_SMALL];
  double opn[KQ_MMAX_SMALL];
#pragma unroll
  for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
    t2p[m] = -2;
    opn[m] = 0.0;
    if (m < M) {
      t2p[m] = a.term2pulse[k * M + m];
      opn[m] = a.op_norm[k * M + m];
      for (int e = 0; e < NN; ++e)
        sm.ops[((size_t)m * NN + e) * BT + tid] = a.ops[((size_t)k * M + m) * NN + e];
    }
  }
  cplx y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = a.state0[(size_t)k * N + i];

  const int n_first = a.backward ? NT - 1 : 0;
  const int n_step = a.backward ? -1 : 1;
  if (a.store) {
    const size_t row = a.backward ? (size_t)NT : 0;
#pragma unroll
    for (int i = 0; i < N; ++i) kq_store(a, (row * K + k) * N + i, y[i]);
  }
  // coefficients of the first step
  double coef[KQ_MMAX_SMALL];
#pragma unroll
  for (int m = 0; m < KQ_MMAX_SMALL; ++m)
    coef[m] = (t2p[m] >= 0) ? a.pulses[(size_t)t2p[m] * NT + n_first] : (t2p[m] == -1 ? 1.0 : 0.0);
  double dtn = a.dt[n_first];

  for (int it = 0, n = n_first; it < NT; ++it, n += n_step) {
    // prefetch the next step's scalars while this step computes
    const int nn = (it + 1 < NT) ? n + n_step : n;
    double coef_next[KQ_MMAX_SMALL];
#pragma unroll
    for (int m = 0; m < KQ_MMAX_SMALL; ++m)
      coef_next[m] = (t2p[m] >= 0) ? a.pulses[(size_t)t2p[m] * NT + nn] : coef[m];
    const double dt_next = a.dt[nn];

    cplx A[NN];
    double x = 0.0;
#pragma unroll
    for (int e = 0; e < NN; ++e) A[e] = c_zero();
#pragma unroll
    for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
      if (m < M) {
        x = fma(fabs(coef[m]), opn[m], x);
#pragma unroll
        for (int e = 0; e < NN; ++e)
          A[e] = c_fma_real(coef[m], sm.ops[((size_t)m * NN + e) * BT + tid], A[e]);
      }
    }
    expmv_small<N, FSEL>(T, A, y, dtn, x * dtn);
    if (a.store) {
      const size_t row = a.backward ? (size_t)n : (size_t)n + 1;
#pragma unroll
      for (int i = 0; i < N; ++i) kq_store(a, (row * K + k) * N + i, y[i]);
    }
#pragma unroll
    for (int m = 0; m < KQ_MMAX_SMALL; ++m) coef[m] = coef_next[m];
    dtn = dt_next;
  }
  if (a.stateT) {
#pragma unroll
    for (int i = 0; i < N; ++i) a.stateT[(size_t)k * N + i] = y[i];
  }
}

// ---------------------------------------------------------------------------
// Fused pulse update + forward step (optimize.py:449-500).
template <int N, int FSEL, bool SECOND, int BTMAX>
__global__ void __launch_bounds__(BTMAX, 1)
k_fwupd_small(const KqSweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // fall-back of a faster sweep kernel queued before this one: run only if it asked for it
  if (a.cond_epoch && *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch) return;
  const KqTables& T = c_kq_tables;
  constexpr int NN = N * N;
  const int BT = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = (BT + 31) >> 5;
  const int k = blockIdx.x * BT + tid;
  const int K = a.K, NT = a.NT, M = a.M, L = a.L;
  const bool valid = k < K;
  const int kk = valid ? k : K - 1;  // padding threads shadow the last objective, weight 0
  const int nblk = gridDim.x;
  const bool writer = (blockIdx.x == 0 && tid == 0);
  SmallSmem<N> sm(smem_raw, M, L, BT, true);

  int t2p[KQ_MMAX_SMALL];
  double opn[KQ_MMAX_SMALL], lam[KQ_MMAX_SMALL];
#pragma unroll
  for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
    t2p[m] = -2;
    opn[m] = 0.0;
    lam[m] = 1.0;
    if (m < M) {
      t2p[m] = a.term2pulse[kk * M + m];
      opn[m] = a.op_norm[kk * M + m];
      if (t2p[m] >= 0) lam[m] = a.lambda_a[t2p[m]];
      for (int e = 0; e < NN; ++e)
        sm.ops[((size_t)m * NN + e) * BT + tid] = a.ops[((size_t)kk * M + m) * NN + e];
    }
  }
  for (int l = 0; l < L; ++l)
    for (int e = 0; e < NN; ++e)
      sm.mu[((size_t)l * NN + e) * BT + tid] = a.mu[((size_t)kk * L + l) * NN + e];

  cplx phi[N], chi[N], dphi[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    phi[i] = a.state0[(size_t)kk * N + i];
    chi[i] = a.X[((size_t)0 * K + kk) * N + i];
    dphi[i] = c_zero();
  }
  const double cnorm = valid ? a.chi_norms[kk] : 0.0;
  if (SECOND && a.store) {
    if (valid) {
#pragma unroll
      for (int i = 0; i < N; ++i) a.store[((size_t)0 * K + k) * N + i] = phi[i];
    }
  }
  double ga[KQ_LMAX];
#pragma unroll
  for (int l = 0; l < KQ_LMAX; ++l) ga[l] = 0.0;
  bool failed = false;

  for (int n = 0; n < NT; ++n) {
    const int par = n & 1;
    // ---- loads that do not depend on the sequential chain -----------------
    cplx chi_next[N], p0_next[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      chi_next[i] = a.X[((size_t)(n + 1) * K + kk) * N + i];
      if (SECOND) p0_next[i] = a.Phi0[((size_t)(n + 1) * K + kk) * N + i];
    }
    const double dtn = a.dt[n];
    const double sig = SECOND ? a.sigma[n] : 0.0;
    double gs[KQ_MMAX_SMALL], ss[KQ_MMAX_SMALL];  // guess pulse, S/lambda per term
#pragma unroll
    for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
      gs[m] = 0.0;
      ss[m] = 0.0;
      if (t2p[m] >= 0) {
        gs[m] = a.pulses[(size_t)t2p[m] * NT + n];
        ss[m] = a.shape[(size_t)t2p[m] * NT + n] / lam[m];
      }
    }
    // ---- Im <chi| mu_l |phi> (+ second-order term), summed over objectives --
    for (int l = 0; l < L; ++l) {
      double val = 0.0, val2 = 0.0;
#pragma unroll
      for (int r = 0; r < N; ++r) {
        cplx w = c_zero();
#pragma unroll
        for (int c = 0; c < N; ++c)
          w = c_fma(sm.mu[((size_t)l * NN + c * N + r) * BT + tid], phi[c], w);
        val += c_im_conj_mul(chi[r], w);
        if (SECOND) val2 += c_im_conj_mul(dphi[r], w);
      }
      val *= cnorm;
      if (SECOND) val = fma(0.5 * sig, valid ? val2 : 0.0, val);
      val = warp_allreduce_sum(val);
      if (lane == 0) sm.red[(par * KQ_LMAX + l) * 32 + warp] = val;
    }
    __syncthreads();
    if (nblk > 1 || a.world > 1) {
      // cross-CTA (and cross-GPU) exchange of the per-CTA sums
      const uint32_t tag = a.tag_base + (uint32_t)n + 1u;
      if (nblk > 1) {
        if (tid < L) {
          double t = 0.0;
          for (int w = 0; w < nwarps; ++w) t += sm.red[(par * KQ_LMAX + tid) * 32 + w];
          slot_store(&a.slots[((size_t)par * nblk + blockIdx.x) * L + tid], t, tag);
        }
        if (warp == 0) {
          for (int l = 0; l < L; ++l) {
            double acc = 0.0;
            for (int c = lane; c < nblk; c += 32)
              acc += slot_wait(&a.slots[((size_t)par * nblk + c) * L + l], tag, failed);
            acc = warp_allreduce_sum(acc);
            if (lane == 0) sm.tot[par * KQ_LMAX + l] = acc;
          }
        }
      } else if (warp == 0) {
        for (int l = 0; l < L; ++l) {
          double acc = 0.0;
          for (int w = lane; w < nwarps; w += 32) acc += sm.red[(par * KQ_LMAX + l) * 32 + w];
          acc = warp_allreduce_sum(acc);
          if (lane == 0) sm.tot[par * KQ_LMAX + l] = acc;
        }
      }
      if (a.world > 1) {
        // second level: one value per GPU, published into every peer's buffer
        // by CTA 0, gathered from the local buffer by every CTA (rank order).
        __syncthreads();
        KqSlot* mine = a.peer_slots[a.rank];
        const size_t goff = KQ_RANK_SLOT_OFFSET;
        if (blockIdx.x == 0 && tid < L * a.world) {
          const int l = tid % L, r = tid / L;
          KqSlot* dst = a.peer_slots[r] + goff + ((size_t)par * a.world + a.rank) * KQ_LMAX + l;
          slot_store(dst, sm.tot[par * KQ_LMAX + l], tag);
        }
        __syncthreads();
        if (warp == 0) {
          for (int l = 0; l < L; ++l) {
            double acc = 0.0;
            if (lane == 0) {
              for (int r = 0; r < a.world; ++r)
                acc += slot_wait(mine + goff + ((size_t)par * a.world + r) * KQ_LMAX + l, tag,
                                 failed);
              sm.tot[par * KQ_LMAX + l] = acc;
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- pulse update and generator under the UPDATED pulse ----------------
    cplx A[NN];
    double x = 0.0;
#pragma unroll
    for (int e = 0; e < NN; ++e) A[e] = c_zero();
#pragma unroll
    for (int m = 0; m < KQ_MMAX_SMALL; ++m) {
      if (m < M) {
        double coef = (t2p[m] == -1) ? 1.0 : 0.0;
        if (t2p[m] >= 0) {
          double d1;
          if (nblk > 1 || a.world > 1) {
            d1 = sm.tot[par * KQ_LMAX + t2p[m]];
          } else {
            const double* rp = &sm.red[(par * KQ_LMAX + t2p[m]) * 32];
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int w = 0;
            for (; w + 3 < nwarps; w += 4) {
              s0 += rp[w];
              s1 += rp[w + 1];
              s2 += rp[w + 2];
              s3 += rp[w + 3];
            }
            for (; w < nwarps; ++w) s0 += rp[w];
            d1 = (s0 + s1) + (s2 + s3);
          }
          // eps + (S/lambda) * Im(...), rounded like optimize.py:474,477
          coef = __dadd_rn(gs[m], __dmul_rn(ss[m], d1));
        }
        x = fma(fabs(coef), opn[m], x);
#pragma unroll
        for (int e = 0; e < NN; ++e)
          A[e] = c_fma_real(coef, sm.ops[((size_t)m * NN + e) * BT + tid], A[e]);
      }
    }
    if (writer) {
      for (int l = 0; l < L; ++l) {
        double d1;
        if (nblk > 1 || a.world > 1) {
          d1 = sm.tot[par * KQ_LMAX + l];
        } else {
          const double* rp = &sm.red[(par * KQ_LMAX + l) * 32];
          double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
          int w = 0;
          for (; w + 3 < nwarps; w += 4) {
            s0 += rp[w];
            s1 += rp[w + 1];
            s2 += rp[w + 2];
            s3 += rp[w + 3];
          }
          for (; w < nwarps; ++w) s0 += rp[w];
          d1 = (s0 + s1) + (s2 + s3);
        }
        const double sl = a.shape[(size_t)l * NT + n] / a.lambda_a[l];
        a.opt_pulses[(size_t)l * NT + n] =
            __dadd_rn(a.pulses[(size_t)l * NT + n], __dmul_rn(sl, d1));
        // (S/lambda) * |d1|^2 * dt accumulated as in optimize.py:475
        ga[l] = __dadd_rn(ga[l], __dmul_rn(__dmul_rn(sl, __dmul_rn(d1, d1)), dtn));
      }
    }
    expmv_small<N, FSEL>(T, A, phi, dtn, x * dtn);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      chi[i] = chi_next[i];
      if (SECOND) dphi[i] = c_sub(phi[i], p0_next[i]);
    }
    if (SECOND && a.store && valid) {
#pragma unroll
      for (int i = 0; i < N; ++i) a.store[((size_t)(n + 1) * K + k) * N + i] = phi[i];
    }
  }
  if (a.stateT && valid) {
#pragma unroll
    for (int i = 0; i < N; ++i) a.stateT[(size_t)k * N + i] = phi[i];
  }
  if (writer) {
    for (int l = 0; l < L; ++l) a.g_a[l] = ga[l];
  }
  if (failed) atomicExch(a.status, (int)-4);
}
