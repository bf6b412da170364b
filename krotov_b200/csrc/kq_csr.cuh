// Row-per-thread kernels for LARGE state vectors (64 < N <= 1024) with SPARSE generator
// terms -- the Liouville-space problems of the reference's notebook 06 (two five-level
// transmons: density matrices 25 x 25, super-operators 625 x 625 with ~60 non-zeros per row;
// /root/reference/docs/notebooks/06_example_3states.ipynb).
//
// One CTA owns one objective, one thread one row of the state vector.  The generator terms,
// their adjoints and mu come as CSR matrices (kq_sparse, built once by the problem compiler);
// the Taylor/Horner recurrence of propagators.expm (see kq_common.cuh) is carried out on the
// vector, y <- v + (h/j) f (sum_m c_m T_m) y, with the state exchanged through a
// double-buffered shared-memory vector and one __syncthreads per Horner step.  The pulse
// update (optimize.py:449-500) sums Im <chi_k| mu_l |phi_k> over the rows of a CTA (shuffles +
// shared memory) and over the objectives through the flag-tagged slots in global memory the
// lane-per-row family uses (co-resident grid, deterministic order).
#pragma once
#include "kq_common.cuh"
#include "kq_small.cuh"   // KqSweepArgs

struct KqCsr {
  const int* row_ptr;         // [n_mat][N+1], relative to the matrix' offset
  const long long* mat_off;   // [n_mat+1] offsets into col / val / col16 / code16
  const int* col;
  const cplx* val;
  // dictionary-coded copy (Liouvillians repeat few distinct values): val = dict[code16]
  const unsigned short* col16;
  const unsigned short* code16;
  const cplx* dict;
  int n_dict;
};

// (T x)[row] for CSR matrix `mat`; x in shared memory
__device__ __forceinline__ cplx csr_row(const KqCsr& s, int mat, int N, int row, const cplx* x) {
  const int* rp = s.row_ptr + (size_t)mat * (N + 1) + row;
  const long long off = s.mat_off[mat];
  const int p0 = rp[0], p1 = rp[1];
  const int* col = s.col + off;
  const cplx* val = s.val + off;
  cplx a0 = c_zero(), a1 = c_zero();
  int p = p0;
  for (; p + 1 < p1; p += 2) {
    a0 = c_fma(val[p], x[col[p]], a0);
    a1 = c_fma(val[p + 1], x[col[p + 1]], a1);
  }
  if (p < p1) a0 = c_fma(val[p], x[col[p]], a0);
  return c_add(a0, a1);
}

// The matrices one CTA works with, staged in shared memory: dictionary | per matrix
// row pointers (int, relative) | packed (column, code) pairs.
struct CsrStage {
  const cplx* dict;
  const int* rp;            // [n_local][N+1], entries index into `cc`
  const unsigned int* cc;   // column | code << 16
};
__device__ __forceinline__ cplx csr_row_staged(const CsrStage& g, int local, int N, int row,
                                               const cplx* x) {
  const int* rp = g.rp + local * (N + 1) + row;
  const int p0 = rp[0], p1 = rp[1];
  cplx a0 = c_zero(), a1 = c_zero();
  int p = p0;
  for (; p + 1 < p1; p += 2) {
    const unsigned int e0 = g.cc[p], e1 = g.cc[p + 1];
    a0 = c_fma(g.dict[e0 >> 16], x[e0 & 0xffffu], a0);
    a1 = c_fma(g.dict[e1 >> 16], x[e1 & 0xffffu], a1);
  }
  if (p < p1) {
    const unsigned int e0 = g.cc[p];
    a0 = c_fma(g.dict[e0 >> 16], x[e0 & 0xffffu], a0);
  }
  return c_add(a0, a1);
}

// matrix numbering of kq_sparse: generator terms | adjoint terms | mu
__device__ __forceinline__ int csr_mat_op(const KqSweepArgs& a, bool adjoint, int k, int m) {
  return (adjoint ? a.K * a.M : 0) + k * a.M + m;
}
__device__ __forceinline__ int csr_mat_mu(const KqSweepArgs& a, int k, int l) {
  return 2 * a.K * a.M + k * a.L + l;
}

// shared: red [2][KQ_LMAX][32] | tot [2][KQ_LMAX] | scoef [M padded] | xb [2][N] cplx
//         | STAGED: dict [n_dict] cplx | rp [(M (+L)) (N+1)] int | cc [nnz] uint
template <int FSEL, bool UPDATE, bool STAGED>
__global__ void __launch_bounds__(1024, 1) k_sweep_csr(const KqSweepArgs a, const KqCsr s) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (UPDATE && a.cond_epoch &&
      *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch)
    return;
  const KqTables& T = c_kq_tables;
  const int K = a.K, N = a.N, NT = a.NT, M = a.M, L = a.L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int nblk = gridDim.x;
  const int kk = (UPDATE ? 0 : a.k_lo) + blockIdx.x;
  const int row = tid;
  const bool act = row < N;
  double* red = reinterpret_cast<double*>(smem_raw);   // [2][KQ_LMAX][32]
  double* tot = red + 2 * KQ_LMAX * 32;                // [2][KQ_LMAX]
  double* scoef = tot + 2 * KQ_LMAX;                   // [M]
  cplx* xb = reinterpret_cast<cplx*>(scoef + ((M + 1) & ~1));   // [2][N]
  const int* t2p = a.term2pulse + (size_t)kk * M;
  const double* opn = a.op_norm + (size_t)kk * M;
  const bool bwd = !UPDATE && a.backward;
  // STAGED: this objective's matrices (generator terms 0..M-1, then mu_0..mu_{L-1} for the
  // update sweep) live in shared memory for the whole sweep
  CsrStage sg;
  if (STAGED) {
    cplx* sdict = xb + 2 * N;
    int* srp = reinterpret_cast<int*>(sdict + s.n_dict);
    const int nloc = UPDATE ? M + L : M;
    unsigned int* scc = reinterpret_cast<unsigned int*>(srp + nloc * (N + 1));
    for (int i = tid; i < s.n_dict; i += blockDim.x) sdict[i] = s.dict[i];
    int base = 0;
    for (int j = 0; j < nloc; ++j) {
      const int mat = j < M ? csr_mat_op(a, bwd, kk, j) : csr_mat_mu(a, kk, j - M);
      const long long off = s.mat_off[mat];
      const int nnz = (int)(s.mat_off[mat + 1] - off);
      const int* rp = s.row_ptr + (size_t)mat * (N + 1);
      for (int i = tid; i <= N; i += blockDim.x) srp[j * (N + 1) + i] = base + rp[i];
      for (int i = tid; i < nnz; i += blockDim.x)
        scc[base + i] = (unsigned int)s.col16[off + i] | ((unsigned int)s.code16[off + i] << 16);
      base += nnz;
    }
    sg.dict = sdict;
    sg.rp = srp;
    sg.cc = scc;
  }

  cplx y = c_zero(), chi = c_zero();
  if (act) {
    y = a.state0[(size_t)kk * N + row];
    if (UPDATE) chi = a.X[(size_t)kk * N + row];
    xb[row] = y;
  }
  const double cnorm = UPDATE ? a.chi_norms[kk] : 0.0;
  if (a.store && !UPDATE && act)
    a.store[(((size_t)(bwd ? NT : 0)) * K + kk) * N + row] = y;
  double ga = 0.0;
  bool failed = false;
  int p = 0;
  __syncthreads();
  for (int it = 0; it < NT; ++it) {
    const int n = bwd ? NT - 1 - it : it;
    const int par = it & 1;
    const double dtn = a.dt[n];
    cplx chi_next = c_zero();
    if (UPDATE) {
      if (act) chi_next = a.X[((size_t)(n + 1) * K + kk) * N + row];
      double pg = 0.0, psl = 0.0;
      if (tid < L) {
        pg = a.pulses[(size_t)tid * NT + n];
        psl = a.shape[(size_t)tid * NT + n] / a.lambda_a[tid];   // optimize.py:474
      }
      // ---- Im <chi| mu_l |phi> ||chi||, summed over the rows and over all objectives
      const cplx* xcur = xb + p * N;
      for (int l = 0; l < L; ++l) {
        double val = 0.0;
        if (act) {
          const cplx w = STAGED ? csr_row_staged(sg, M + l, N, row, xcur)
                                : csr_row(s, csr_mat_mu(a, kk, l), N, row, xcur);
          val = c_im_conj_mul(chi, w) * cnorm;
        }
        val = warp_allreduce_sum(val);
        if (lane == 0) red[(par * KQ_LMAX + l) * 32 + warp] = val;
      }
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
      __syncthreads();
      // updated pulse values (optimize.py:471-477)
      if (tid < L) {
        const double d1 = tot[par * KQ_LMAX + tid];
        const double eps_new = __dadd_rn(pg, __dmul_rn(psl, d1));
        red[par * KQ_LMAX * 32 + tid] = eps_new;   // red[par][0][0..L) is free again
        if (blockIdx.x == 0) {
          a.opt_pulses[(size_t)tid * NT + n] = eps_new;
          ga = __dadd_rn(ga, __dmul_rn(__dmul_rn(psl, __dmul_rn(d1, d1)), dtn));
        }
      }
      __syncthreads();
      if (tid < M) {
        const int l = t2p[tid];
        scoef[tid] = (l == -1) ? 1.0 : (l >= 0 ? red[par * KQ_LMAX * 32 + l] : 0.0);
      }
    } else {
      if (tid < M) {
        const int l = t2p[tid];
        scoef[tid] = (l == -1) ? 1.0 : (l >= 0 ? a.pulses[(size_t)l * NT + n] : 0.0);
      }
    }
    __syncthreads();
    double x = 0.0;
    for (int m = 0; m < M; ++m) x = fma(fabs(scoef[m]), opn[m], x);
    x *= dtn;
    int sc, mdeg;
    taylor_plan_fine(T, x, sc, mdeg);   // quarter-binade degrees: every term is a pass over the matrix
    const double h = (sc == 1) ? dtn : dtn / (double)sc;
    for (int rep = 0; rep < sc; ++rep) {
      const cplx v = y;
      for (int j = mdeg; j >= 1; --j) {
        const double cj = h * T.inv[j];
        const cplx* xcur = xb + p * N;
        if (act) {
          cplx w = c_zero();
          for (int m = 0; m < M; ++m) {
            const double cm = scoef[m];
            if (cm != 0.0)
              w = c_fma_real(cm, STAGED ? csr_row_staged(sg, m, N, row, xcur)
                                        : csr_row(s, csr_mat_op(a, bwd, kk, m), N, row, xcur), w);
          }
          y = c_fma_real(cj, apply_f<FSEL>(w), v);
          xb[(p ^ 1) * N + row] = y;
        }
        __syncthreads();
        p ^= 1;
      }
    }
    if (UPDATE) chi = chi_next;
    if (a.store && !UPDATE && act)
      a.store[(((size_t)(bwd ? n : n + 1)) * K + kk) * N + row] = y;
  }
  if (a.stateT && act) a.stateT[(size_t)kk * N + row] = y;
  if (UPDATE && blockIdx.x == 0 && tid < L) a.g_a[tid] = ga;
  if (UPDATE && failed) atomicExch(a.status, (int)-4);
}
