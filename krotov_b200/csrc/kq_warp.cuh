// Lane-per-row kernels for 5 <= N <= 64 (and small N with many terms).
//
// A group of R = 2^ceil(log2 N) lanes (R = 32 and RPL = 2 rows per lane for
// 32 < N <= 64) owns one objective; a warp holds G = 32/R objectives.  The
// assembled generator A, the generator terms and mu sit in shared memory in
// column-major order so that consecutive lanes (= consecutive rows) read
// consecutive 16-byte words: conflict-free LDS.128, coalesced global fills.
// The state vector is exchanged through a double-buffered shared-memory
// vector with one __syncwarp per Horner step; objectives never span warps.
// One template covers the plain propagation sweeps (UPDATE = false:
// optimize.py:806-886 of the reference) and the fused update + forward sweep
// (UPDATE = true: optimize.py:449-500).
#pragma once
#include "kq_common.cuh"
#include "kq_small.cuh"  // KqSweepArgs

#include "kq_warp_geom.cuh"

// w = sum_c Mcol[c*N + row] * x[c]  (column-major matrix, unrolled by 4)
__device__ __forceinline__ cplx matvec_row(const cplx* __restrict__ Mcol,
                                           const cplx* __restrict__ x, int N, int row) {
  cplx w0 = c_zero(), w1 = c_zero(), w2 = c_zero(), w3 = c_zero();
  const cplx* m = Mcol + row;
  int c = 0;
  for (; c + 3 < N; c += 4) {
    const cplx a0 = m[(size_t)c * N], a1 = m[(size_t)(c + 1) * N];
    const cplx a2 = m[(size_t)(c + 2) * N], a3 = m[(size_t)(c + 3) * N];
    const cplx x0 = x[c], x1 = x[c + 1], x2 = x[c + 2], x3 = x[c + 3];
    w0 = c_fma(a0, x0, w0);
    w1 = c_fma(a1, x1, w1);
    w2 = c_fma(a2, x2, w2);
    w3 = c_fma(a3, x3, w3);
  }
  for (; c < N; ++c) w0 = c_fma(m[(size_t)c * N], x[c], w0);
  return c_add(c_add(w0, w1), c_add(w2, w3));
}

// MODE 0: generic (terms / mu may stay in global memory).
// MODE 1: generator terms and mu resident in shared memory (pointers are then
//         known to be shared-space: LDS instead of generic loads).
// MODE 8/16/32: as 1, and each lane keeps its row of the assembled generator
//         A in registers (N <= MODE, RPL = 1): a Horner step is then one
//         broadcast LDS.128 of the state element + 4 DFMA per column, with no
//         shared-memory round trip for A.
template <int RPL, int FSEL, bool SECOND, bool UPDATE, int MODE>
__global__ void __launch_bounds__(MODE >= 32 ? 128 : (MODE >= 16 ? 256 : 512), 1)
k_sweep_warp(const KqSweepArgs a, const KqWarpGeom g) {
  constexpr bool ALLSM = MODE >= 1;
  constexpr int AREG = (MODE >= 8) ? MODE : 0;
  static_assert(AREG == 0 || RPL == 1, "register rows need one row per lane");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // fall-back of a faster update sweep queued before this one: run only if it asked for it
  if (UPDATE && a.cond_epoch &&
      *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch)
    return;
  const KqTables& T = c_kq_tables;
  const int K = a.K, N = a.N, NT = a.NT, M = a.M, L = a.L, NN = N * N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int R = g.R, G = g.G;
  const int grp = lane / R, lig = lane - grp * R;
  // Task index.  Plain sweeps: one task per objective.  Time-parallel
  // propagation (seg_pass 1/2, UPDATE = false): one task per (objective,
  // segment[, basis vector]) -- pass 1 sends the N basis vectors through every
  // segment (segment propagators), pass 2 propagates every segment from its
  // boundary state and stores all states (see k_prop_spec in kq_spec.cuh).
  const int seg_pass = UPDATE ? 0 : a.seg_pass;
  const int n_task = UPDATE ? K
                            : (seg_pass ? a.k_cnt * a.seg_count * (seg_pass == 1 ? N : 1)
                                        : a.k_cnt);
  int task = (blockIdx.x * nwarps + warp) * G + grp;
  const bool valid = task < n_task;
  if (!valid) task = n_task - 1;
  int seg = 0, vec = 0;
  if (seg_pass) {
    const int rest = task / a.k_cnt;
    task -= rest * a.k_cnt;
    seg = rest % a.seg_count;
    vec = rest / a.seg_count;
  }
  const int ob = (UPDATE ? 0 : a.k_lo) + task;
  const int kk = ob;
  const int nblk = gridDim.x;

  // ---- shared memory carve-up ------------------------------------------
  double* red = reinterpret_cast<double*>(smem_raw);  // [2][KQ_LMAX][32]
  double* tot = red + 2 * KQ_LMAX * 32;               // [2][KQ_LMAX]
  cplx* objbase = reinterpret_cast<cplx*>(tot + 2 * KQ_LMAX) +
                  (size_t)(warp * G + grp) * g.obj_stride;
  // state buffers padded (pads stay zero): to the register-row capacity in
  // MODE >= 8 so that the matvec is branch-free, else to a multiple of 4
  const int NP4 = (AREG > 0) ? AREG : ((N + 3) & ~3);
  cplx* sA = objbase;                   // [N*N] column-major
  cplx* xb = sA + NN;                   // [2][NP4], pads stay zero
  double* scoef = reinterpret_cast<double*>(xb + 2 * NP4);  // [M] current coefficients
  cplx* sterms = reinterpret_cast<cplx*>(scoef + ((M + 1) & ~1));
  cplx* smu = sterms + ((ALLSM || g.terms_in_smem) ? (size_t)M * NN : 0);
  const cplx* gterms = a.ops + (size_t)kk * M * NN;
  const cplx* gmu = UPDATE ? a.mu + (size_t)kk * L * NN : nullptr;
  if (ALLSM || g.terms_in_smem) {
    for (int e = lig; e < M * NN; e += R) sterms[e] = gterms[e];
  }
  if (UPDATE && (ALLSM || g.mu_in_smem)) {
    for (int e = lig; e < L * NN; e += R) smu[e] = gmu[e];
  }
  const cplx* terms = ALLSM ? sterms : (g.terms_in_smem ? sterms : gterms);
  const cplx* mu = ALLSM ? smu : ((UPDATE && g.mu_in_smem) ? smu : gmu);
  const int* t2p = a.term2pulse + (size_t)kk * M;
  const double* opn = a.op_norm + (size_t)kk * M;

  // ---- per-lane rows -------------------------------------------------------
  int row[RPL];
  bool act[RPL];
  cplx y[RPL], chi[RPL], dphi[RPL];
#pragma unroll
  for (int q = 0; q < RPL; ++q) {
    row[q] = lig + R * q;
    act[q] = row[q] < N;
    y[q] = c_zero();
    if (act[q]) {
      if (seg_pass == 1)
        y[q] = c_make(row[q] == vec ? 1.0 : 0.0, 0.0);
      else if (seg_pass == 2)
        y[q] = a.seg_B[((size_t)seg * K + kk) * N + row[q]];
      else
        y[q] = a.state0[(size_t)kk * N + row[q]];
    }
    chi[q] = c_zero();
    dphi[q] = c_zero();
    if (UPDATE && act[q]) chi[q] = a.X[((size_t)0 * K + kk) * N + row[q]];
    if (act[q]) xb[row[q]] = y[q];
  }
  for (int e = N + lig; e < NP4; e += R) {
    xb[e] = c_zero();
    xb[NP4 + e] = c_zero();
  }
  cplx arow[AREG > 0 ? AREG : 1];
  const double cnorm = (UPDATE && valid) ? a.chi_norms[kk] : 0.0;
  const bool bwd = !UPDATE && a.backward;
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
  if (a.store && valid && (!UPDATE || SECOND) && seg == 0 && seg_pass != 1) {
    const size_t r0 = (!UPDATE && a.backward) ? (size_t)NT : 0;
#pragma unroll
    for (int q = 0; q < RPL; ++q)
      if (act[q]) {
        if (UPDATE) a.store[(r0 * K + ob) * N + row[q]] = y[q];
        else kq_store(a, (r0 * K + ob) * N + row[q], y[q]);
      }
  }
  double ga = 0.0;       // lane l of warp 0 in CTA 0 accumulates g_a[l]
  const bool solo = UPDATE && (nwarps == 1) && (nblk == 1) && (a.world == 1);
  bool failed = false;
  int p = 0;  // xb[p] holds the current state
  __syncwarp();

  for (int it = 0, n = n_first; it < n_count; ++it, n += n_step) {
    const int par = it & 1;
    const double dtn = a.dt[n];
    cplx chi_next[RPL], p0_next[RPL];
    if (UPDATE) {
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        chi_next[q] = c_zero();
        p0_next[q] = c_zero();
        if (act[q]) {
          chi_next[q] = a.X[((size_t)(n + 1) * K + kk) * N + row[q]];
          if (SECOND) p0_next[q] = a.Phi0[((size_t)(n + 1) * K + kk) * N + row[q]];
        }
      }
      const double sig = SECOND ? a.sigma[n] : 0.0;
      // per-pulse scalars of this step, loaded before the reduction so that no
      // global load or division sits on the sequential chain
      double pg = 0.0, psl = 0.0;
      if (lane < L) {
        pg = a.pulses[(size_t)lane * NT + n];
        psl = a.shape[(size_t)lane * NT + n] / a.lambda_a[lane];   // optimize.py:474
      }
      // ---- Im <chi| mu_l |phi>, summed over all objectives -----------------
      const cplx* xcur = xb + p * NP4;
      for (int l = 0; l < L; ++l) {
        double val = 0.0, val2 = 0.0;
        const cplx* Ml = mu + (size_t)l * NN;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          if (act[q]) {
            const cplx w = matvec_row(Ml, xcur, N, row[q]);
            val += c_im_conj_mul(chi[q], w);
            if (SECOND) val2 += c_im_conj_mul(dphi[q], w);
          }
        }
        val *= cnorm;
        if (SECOND) val = fma(0.5 * sig, valid ? val2 : 0.0, val);
        val = warp_allreduce_sum(val);
        if (solo) {
          if (lane == 0) tot[par * KQ_LMAX + l] = val;
        } else if (lane == 0) {
          red[(par * KQ_LMAX + l) * 32 + warp] = val;
        }
      }
      if (solo) {
        __syncwarp();
      } else {
      __syncthreads();
      if (warp == 0) {
        const uint32_t tag = a.tag_base + (uint32_t)n + 1u;
        for (int l = 0; l < L; ++l) {
          double acc = (lane < nwarps) ? red[(par * KQ_LMAX + l) * 32 + lane] : 0.0;
          acc = warp_allreduce_sum(acc);
          if (nblk > 1) {
            if (lane == 0)
              slot_store(&a.slots[((size_t)par * nblk + blockIdx.x) * L + l], acc, tag);
            double g2 = 0.0;
            for (int c = lane; c < nblk; c += 32)
              g2 += slot_wait(&a.slots[((size_t)par * nblk + c) * L + l], tag, failed);
            acc = warp_allreduce_sum(g2);
          }
          if (lane == 0) tot[par * KQ_LMAX + l] = acc;
        }
      }
      if (a.world > 1) {
        const uint32_t tag = a.tag_base + (uint32_t)n + 1u;
        __syncthreads();
        KqSlot* mine = a.peer_slots[a.rank];
        const size_t goff = KQ_RANK_SLOT_OFFSET;
        if (blockIdx.x == 0 && tid < L * a.world) {
          const int l = tid % L, r = tid / L;
          slot_store(a.peer_slots[r] + goff + ((size_t)par * a.world + a.rank) * KQ_LMAX + l,
                     tot[par * KQ_LMAX + l], tag);
        }
        __syncthreads();
        if (tid == 0) {
          for (int l = 0; l < L; ++l) {
            double acc = 0.0;
            for (int r = 0; r < a.world; ++r)
              acc += slot_wait(mine + goff + ((size_t)par * a.world + r) * KQ_LMAX + l, tag, failed);
            tot[par * KQ_LMAX + l] = acc;
          }
        }
      }
      __syncthreads();
      }
      // updated pulse values (optimize.py:471-477); every warp computes them
      // redundantly for its own coefficients, warp 0 of CTA 0 records them
      double eps_new = 0.0;
      if (lane < L) {
        const double d1 = tot[par * KQ_LMAX + lane];
        eps_new = __dadd_rn(pg, __dmul_rn(psl, d1));
        if (blockIdx.x == 0 && warp == 0) {
          a.opt_pulses[(size_t)lane * NT + n] = eps_new;
          ga = __dadd_rn(ga, __dmul_rn(__dmul_rn(psl, __dmul_rn(d1, d1)), dtn));
        }
      }
      // coefficients of this step's generator under the updated pulses
      for (int m = lig; m < M; m += R) {
        const int l = t2p[m];
        scoef[m] = (l == -1) ? 1.0 : 0.0;
      }
      for (int l = 0; l < L; ++l) {
        const double e = __shfl_sync(0xffffffffu, eps_new, l);
        for (int m = lig; m < M; m += R)
          if (t2p[m] == l) scoef[m] = e;
      }
    }
    if (!UPDATE) {
      // ---- coefficients of this step's generator --------------------------
      for (int m = lig; m < M; m += R) {
        const int l = t2p[m];
        double c = (l == -1) ? 1.0 : 0.0;
        if (l >= 0) c = a.pulses[(size_t)l * NT + n];
        scoef[m] = c;
      }
    }
    __syncwarp();
    double x = 0.0;
    for (int m = 0; m < M; ++m) x = fma(fabs(scoef[m]), opn[m], x);
    x *= dtn;
    if (AREG > 0) {
      // assemble this lane's row of A = sum_m coef_m T_m in registers
#pragma unroll
      for (int c = 0; c < AREG; ++c) arow[c] = c_zero();
      if (act[0]) {
        for (int m = 0; m < M; ++m) {
          const double cm = scoef[m];
          const cplx* tm = terms + (size_t)m * NN + row[0];
#pragma unroll
          for (int c = 0; c < AREG; ++c) {
            const cplx t = (c < N) ? tm[(size_t)c * N] : c_zero();
            arow[c] = c_fma_real(cm, t, arow[c]);
          }
        }
      }
    } else {
      // assemble A = sum_m coef_m T_m (all lanes of the group, strided elements)
      for (int e = lig; e < NN; e += R) {
        cplx acc = c_zero();
        for (int m = 0; m < M; ++m) acc = c_fma_real(scoef[m], terms[(size_t)m * NN + e], acc);
        sA[e] = acc;
      }
    }
    // warp-uniform Taylor plan from the largest bound in the warp
    {
      const int hi = __reduce_max_sync(0xffffffffu, __double2hiint(x));
      x = __hiloint2double(hi, (int)0xffffffff);
    }
    int s, mdeg;
    taylor_plan_fine(T, x, s, mdeg);   // quarter-binade degrees
    const double h = (s == 1) ? dtn : dtn / (double)s;
    __syncwarp();
    for (int rep = 0; rep < s; ++rep) {
      cplx v[RPL];
#pragma unroll
      for (int q = 0; q < RPL; ++q) v[q] = y[q];
      for (int j = mdeg; j >= 1; --j) {
        const double cj = h * T.inv[j];
        const cplx* xcur = xb + p * NP4;
        cplx* xnext = xb + (p ^ 1) * NP4;
        if (AREG > 0) {
          if (act[0]) {
            // branch-free: arow is zero beyond N and the state pads are zero;
            // all AREG broadcast loads are issued before the FMAs
            cplx xv[AREG > 0 ? AREG : 1];
#pragma unroll
            for (int c = 0; c < AREG; ++c) xv[c] = xcur[c];
            cplx w0 = c_zero(), w1 = c_zero(), w2 = c_zero(), w3 = c_zero();
#pragma unroll
            for (int cb = 0; cb < AREG; cb += 4) {
              w0 = c_fma(arow[cb], xv[cb], w0);
              w1 = c_fma(arow[cb + 1], xv[cb + 1], w1);
              w2 = c_fma(arow[cb + 2], xv[cb + 2], w2);
              w3 = c_fma(arow[cb + 3], xv[cb + 3], w3);
            }
            const cplx w = c_add(c_add(w0, w1), c_add(w2, w3));
            y[0] = c_fma_real(cj, apply_f<FSEL>(w), v[0]);
            xnext[row[0]] = y[0];
          }
        } else {
#pragma unroll
          for (int q = 0; q < RPL; ++q) {
            if (act[q]) {
              const cplx w = matvec_row(sA, xcur, N, row[q]);
              y[q] = c_fma_real(cj, apply_f<FSEL>(w), v[q]);
              xnext[row[q]] = y[q];
            }
          }
        }
        __syncwarp();
        p ^= 1;
      }
    }
    if (UPDATE) {
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        chi[q] = chi_next[q];
        if (SECOND) dphi[q] = c_sub(y[q], p0_next[q]);
      }
    }
    if (a.store && valid && (!UPDATE || SECOND) && seg_pass != 1) {
      const size_t r1 = (!UPDATE && a.backward) ? (size_t)n : (size_t)n + 1;
#pragma unroll
      for (int q = 0; q < RPL; ++q)
        if (act[q]) {
          if (UPDATE) a.store[(r1 * K + ob) * N + row[q]] = y[q];
          else kq_store(a, (r1 * K + ob) * N + row[q], y[q]);
        }
    }
  }
  if (seg_pass == 1) {
    // column `vec` of the segment propagator (column-major)
    if (valid) {
#pragma unroll
      for (int q = 0; q < RPL; ++q)
        if (act[q]) a.seg_P[((size_t)seg * K + ob) * NN + (size_t)vec * N + row[q]] = y[q];
    }
  } else if (a.stateT && valid && (seg_pass == 0 || seg == a.seg_count - 1)) {
#pragma unroll
    for (int q = 0; q < RPL; ++q)
      if (act[q]) a.stateT[(size_t)ob * N + row[q]] = y[q];
  }
  if (UPDATE && blockIdx.x == 0 && warp == 0 && lane < L) a.g_a[lane] = ga;
  if (UPDATE && failed) atomicExch(a.status, (int)-4);
}
