/* TEST INFRASTRUCTURE ONLY -- C restatement of the Krotov sweep hot path.
 *
 * Same algorithm as oracle/krotov_oracle.py (which is pinned against golden
 * vectors from the unmodified reference), written in plain C99 with OpenMP
 * over the objectives so that bench.py can report a multi-threaded CPU
 * baseline and the tests can check large configurations quickly.  Only
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
 *
 * Reference lines followed (paths relative to /root/reference):
 *   kqo_expm_apply   propagators.py:79-122 (A = f*Op0 + sum f*c_m*Op_m,
 *                    expm(A*dt) @ state); the matrix exponential restates the
 *                    published scaling-and-squaring Pade algorithm that
 *                    scipy.linalg.expm implements (Higham 2005 / Al-Mohy &
 *                    Higham 2009: orders 3,5,7,9,13 chosen from ||A||_1)
 *   kqo_backward     optimize.py:849-886
 *   kqo_forward      optimize.py:806-846
 *   kqo_update_sweep optimize.py:449-500 (mu: mu.py:123-140 precomputed by the
 *                    caller, overlap: second_order.py:69-83, update :471-477)
 *
 * Layouts: ops[K][M][N][N] ROW-major complex (re,im), term2pulse[K][M]
 * (-1 drift, l >= 0 pulse, -2 padding), pulses[L][NT], states [K][N],
 * stores [K][NT+1][N] (objective-major, like the reference's storage arrays).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

static double norm1(const cplx* A, int n) {
  double best = 0.0;
  for (int c = 0; c < n; ++c) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += cabs(A[r * n + c]);
    if (s > best) best = s;
  }
  return best;
}

static void matmul(const cplx* A, const cplx* B, cplx* C, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      cplx s = 0;
      for (int k = 0; k < n; ++k) s += A[i * n + k] * B[k * n + j];
      C[i * n + j] = s;
    }
}

/* solve P X = Q in place (X returned in Q), partial pivoting */
static int lu_solve(cplx* P, cplx* Q, int n) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = cabs(P[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (cabs(P[r * n + c]) > best) { best = cabs(P[r * n + c]); piv = r; }
    if (best == 0.0) return -1;
    if (piv != c)
      for (int j = 0; j < n; ++j) {
        cplx t = P[c * n + j]; P[c * n + j] = P[piv * n + j]; P[piv * n + j] = t;
        t = Q[c * n + j]; Q[c * n + j] = Q[piv * n + j]; Q[piv * n + j] = t;
      }
    for (int r = c + 1; r < n; ++r) {
      const cplx f = P[r * n + c] / P[c * n + c];
      if (f == 0) continue;
      for (int j = c; j < n; ++j) P[r * n + j] -= f * P[c * n + j];
      for (int j = 0; j < n; ++j) Q[r * n + j] -= f * Q[c * n + j];
    }
  }
  for (int c = n - 1; c >= 0; --c)
    for (int j = 0; j < n; ++j) {
      cplx s = Q[c * n + j];
      for (int k = c + 1; k < n; ++k) s -= P[c * n + k] * Q[k * n + j];
      Q[c * n + j] = s / P[c * n + c];
    }
  return 0;
}

static const double PADE3[] = {120., 60., 12., 1.};
static const double PADE5[] = {30240., 15120., 3360., 420., 30., 1.};
static const double PADE7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
static const double PADE9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240.,
                               2162160., 110880., 3960., 90., 1.};
static const double PADE13[] = {64764752532480000., 32382376266240000., 7771770303897600.,
                                1187353796428800.,  129060195264000.,   10559470521600.,
                                670442572800.,      33522128640.,       1323241920.,
                                40840800.,          960960.,            16380.,
                                182.,               1.};

/* E = expm(A), scratch w of 7*n*n complex */
static int expm_pade(const cplx* Ain, cplx* E, int n, cplx* w) {
  const int nn = n * n;
  cplx *A = w, *A2 = w + nn, *A4 = w + 2 * nn, *A6 = w + 3 * nn, *U = w + 4 * nn,
       *V = w + 5 * nn, *T = w + 6 * nn;
  memcpy(A, Ain, nn * sizeof(cplx));
  const double nA = norm1(A, n);
  const double th[] = {1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1,
                       2.097847961257068e0, 5.371920351148152e0};
  int order = 13, s = 0;
  if (nA <= th[0]) order = 3;
  else if (nA <= th[1]) order = 5;
  else if (nA <= th[2]) order = 7;
  else if (nA <= th[3]) order = 9;
  else if (nA > th[4]) {
    s = (int)ceil(log2(nA / th[4]));
    if (s < 0) s = 0;
    const double sc = ldexp(1.0, -s);
    for (int i = 0; i < nn; ++i) A[i] *= sc;
  }
  matmul(A, A, A2, n);
  if (order <= 9) {
    const double* b = order == 3 ? PADE3 : order == 5 ? PADE5 : order == 7 ? PADE7 : PADE9;
    /* U = A * sum_{odd} b_k A^(k-1), V = sum_{even} b_k A^k, powers of A2 */
    cplx* P = A4;  /* running power of A2 */
    for (int i = 0; i < nn; ++i) { U[i] = 0; V[i] = 0; P[i] = 0; }
    for (int i = 0; i < n; ++i) { P[i * n + i] = 1; U[i * n + i] = b[1]; V[i * n + i] = b[0]; }
    for (int k = 1; 2 * k <= order; ++k) {
      matmul(P, A2, T, n);
      memcpy(P, T, nn * sizeof(cplx));
      for (int i = 0; i < nn; ++i) {
        U[i] += b[2 * k + 1] * P[i];
        V[i] += b[2 * k] * P[i];
      }
    }
    matmul(A, U, T, n);
    memcpy(U, T, nn * sizeof(cplx));
  } else {
    const double* b = PADE13;
    matmul(A2, A2, A4, n);
    matmul(A4, A2, A6, n);
    for (int i = 0; i < nn; ++i) T[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    matmul(A6, T, U, n);
    for (int i = 0; i < nn; ++i) U[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i];
    for (int i = 0; i < n; ++i) U[i * n + i] += b[1];
    matmul(A, U, T, n);
    memcpy(U, T, nn * sizeof(cplx));
    for (int i = 0; i < nn; ++i) T[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    matmul(A6, T, V, n);
    for (int i = 0; i < nn; ++i) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i];
    for (int i = 0; i < n; ++i) V[i * n + i] += b[0];
  }
  /* E = (V - U)^-1 (V + U) */
  for (int i = 0; i < nn; ++i) { T[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
  if (lu_solve(T, E, n)) return -1;
  for (int q = 0; q < s; ++q) {
    matmul(E, E, T, n);
    memcpy(E, T, nn * sizeof(cplx));
  }
  return 0;
}

/* state <- expm((f * sum_m c_m Op_m) dt) state  (propagators.py:94-117) */
static int expm_apply(const cplx* ops, const int* t2p, int M, int n, const double* pulses,
                      int NT, int step, double dt, cplx f, cplx* state, cplx* w) {
  const int nn = n * n;
  cplx* A = w;         /* nn */
  cplx* E = w + nn;    /* nn */
  cplx* tmp = w + 2 * nn;  /* n */
  for (int i = 0; i < nn; ++i) A[i] = 0;
  for (int m = 0; m < M; ++m) {
    double c;
    if (t2p[m] == -1) c = 1.0;
    else if (t2p[m] >= 0) c = pulses[(size_t)t2p[m] * NT + step];
    else continue;
    const cplx fc = f * c;
    for (int i = 0; i < nn; ++i) A[i] += fc * ops[(size_t)m * nn + i];
  }
  for (int i = 0; i < nn; ++i) A[i] *= dt;
  if (expm_pade(A, E, n, w + 2 * nn + n)) return -1;
  for (int r = 0; r < n; ++r) {
    cplx s = 0;
    for (int c = 0; c < n; ++c) s += E[r * n + c] * state[c];
    tmp[r] = s;
  }
  memcpy(state, tmp, n * sizeof(cplx));
  return 0;
}

static size_t work_size(int n) { return (size_t)9 * n * n + n; }

/* optimize.py:806-846 (store may be NULL) */
int kqo_forward(int K, int N, int NT, int L, int M, const cplx* ops, const int* t2p,
                const double* dt, const double* pulses, const cplx* psi0, cplx* phiT,
                cplx* store, int is_super, int nthreads) {
  int err = 0;
  (void)L;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int k = 0; k < K; ++k) {
    cplx* w = (cplx*)malloc(work_size(N) * sizeof(cplx));
    cplx* st = (cplx*)malloc(N * sizeof(cplx));
    memcpy(st, psi0 + (size_t)k * N, N * sizeof(cplx));
    if (store) memcpy(store + ((size_t)k * (NT + 1)) * N, st, N * sizeof(cplx));
    const cplx f = is_super ? 1.0 : -I;
    for (int n = 0; n < NT; ++n) {
      if (expm_apply(ops + (size_t)k * M * N * N, t2p + (size_t)k * M, M, N, pulses, NT, n,
                     dt[n], f, st, w)) err = 1;
      if (store) memcpy(store + ((size_t)k * (NT + 1) + n + 1) * N, st, N * sizeof(cplx));
    }
    memcpy(phiT + (size_t)k * N, st, N * sizeof(cplx));
    free(w);
    free(st);
  }
  return err;
}

/* optimize.py:849-886: ops_adj = element-wise adjoint operators */
int kqo_backward(int K, int N, int NT, int L, int M, const cplx* ops_adj, const int* t2p,
                 const double* dt, const double* pulses, const cplx* chiT, cplx* X,
                 int is_super, int nthreads) {
  int err = 0;
  (void)L;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int k = 0; k < K; ++k) {
    cplx* w = (cplx*)malloc(work_size(N) * sizeof(cplx));
    cplx* st = (cplx*)malloc(N * sizeof(cplx));
    memcpy(st, chiT + (size_t)k * N, N * sizeof(cplx));
    memcpy(X + ((size_t)k * (NT + 1) + NT) * N, st, N * sizeof(cplx));
    const cplx f = is_super ? 1.0 : I;   /* conj(-i) for backwards=True */
    for (int n = NT - 1; n >= 0; --n) {
      if (expm_apply(ops_adj + (size_t)k * M * N * N, t2p + (size_t)k * M, M, N, pulses, NT,
                     n, dt[n], f, st, w)) err = 1;
      memcpy(X + ((size_t)k * (NT + 1) + n) * N, st, N * sizeof(cplx));
    }
    free(w);
    free(st);
  }
  return err;
}

/* optimize.py:449-500, first order; mu[K][L][N][N] row-major */
int kqo_update_sweep(int K, int N, int NT, int L, int M, const cplx* ops, const cplx* mu,
                     const int* t2p, const double* dt, const double* shape,
                     const double* lambda_a, const double* guess, double* opt, const cplx* X,
                     const double* chi_norms, const cplx* psi0, cplx* phiT, double* g_a,
                     int is_super, int nthreads) {
  int err = 0;
  cplx* phi = (cplx*)malloc((size_t)K * N * sizeof(cplx));
  cplx* part = (cplx*)malloc((size_t)K * L * sizeof(cplx));
  cplx** ws = (cplx**)malloc(nthreads * sizeof(cplx*));
  for (int t = 0; t < nthreads; ++t) ws[t] = (cplx*)malloc(work_size(N) * sizeof(cplx));
  memcpy(phi, psi0, (size_t)K * N * sizeof(cplx));
  memcpy(opt, guess, (size_t)L * NT * sizeof(double));
  for (int l = 0; l < L; ++l) g_a[l] = 0.0;
  const cplx f = is_super ? 1.0 : -I;
#pragma omp parallel num_threads(nthreads)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    for (int n = 0; n < NT; ++n) {
#pragma omp for schedule(static)
      for (int k = 0; k < K; ++k)
        for (int l = 0; l < L; ++l) {
          const cplx* Mu = mu + ((size_t)k * L + l) * N * N;
          const cplx* chi = X + ((size_t)k * (NT + 1) + n) * N;
          cplx ov = 0;
          for (int r = 0; r < N; ++r) {
            cplx s = 0;
            for (int c = 0; c < N; ++c) s += Mu[r * N + c] * phi[(size_t)k * N + c];
            ov += conj(chi[r]) * s;
          }
          part[(size_t)k * L + l] = ov * chi_norms[k];
        }
#pragma omp single
      {
        for (int l = 0; l < L; ++l) {
          cplx acc = 0;
          for (int k = 0; k < K; ++k) acc += part[(size_t)k * L + l];
          const double d1 = cimag(acc);
          const double sl = shape[(size_t)l * NT + n] / lambda_a[l];
          g_a[l] += sl * fabs(d1) * fabs(d1) * dt[n];
          opt[(size_t)l * NT + n] += sl * d1;
        }
      }
#pragma omp for schedule(static)
      for (int k = 0; k < K; ++k)
        if (expm_apply(ops + (size_t)k * M * N * N, t2p + (size_t)k * M, M, N, opt, NT, n,
                       dt[n], f, phi + (size_t)k * N, ws[tid])) err = 1;
    }
  }
  memcpy(phiT, phi, (size_t)K * N * sizeof(cplx));
  for (int t = 0; t < nthreads; ++t) free(ws[t]);
  free(ws);
  free(part);
  free(phi);
  return err;
}

int kqo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
