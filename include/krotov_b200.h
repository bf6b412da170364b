/* krotov_b200.h -- C ABI of libkrotov_b200.so (sm_100a).
 *
 * Drop-in boundary for the Krotov sweep hot path of qucontrol/krotov.  The
 * reference has no FFI: its boundary is the keyword plugin API of
 * krotov.optimize_pulses (src/krotov/optimize.py:33-55), whose per-time-step
 * Python work this library replaces:
 *
 *   kq_propagate_forward     <- _forward_propagation        optimize.py:806-846
 *   kq_sweep_backward        <- _backward_propagation       optimize.py:849-886
 *   kq_sweep_forward_update  <- update + fw-step loop       optimize.py:449-500
 *                               (mu: mu.py:74-140, overlap: second_order.py:69-83,
 *                                propagators.expm: propagators.py:79-122,
 *                                plug_in_pulse_values: conversions.py:288-330)
 *   kq_chi_boundary          <- chi_constructor + norm      optimize.py:404-410,
 *                               functionals.py:177-197,225-253,293-317,389-437
 *   kq_overlaps              <- tau_vals                    optimize.py:316-322,503-508
 *
 * Conventions
 *  - All pointers are DEVICE pointers owned by the caller (torch tensors on
 *    the Python side); the library never allocates or frees user data.  The
 *    only exception is kq_comm setup, which takes host arrays of device
 *    pointers.
 *  - complex128 = two consecutive doubles (re, im) ("kq_c128").
 *  - Every function is asynchronous with respect to `stream` (a cudaStream_t
 *    passed as void*), returns 0 on success and a negative kq_status on
 *    error; kq_last_error() returns a thread-local message.  No exception
 *    crosses the ABI.  No hidden global state besides that message and
 *    per-device constant tables.
 *  - States are vectors of length N: kets, or column-stacked density
 *    matrices when `is_super` is set.
 *
 * Layouts (row = slowest index first)
 *   ops, ops_adj : [K][M][N*N], matrix element (r,c) at c*N + r (column-major)
 *                  term 0..M-1 of objective k's generator; ops_adj holds the
 *                  element-wise adjoint (backward generator).
 *   mu           : [K][L][N*N] column-major, dH/d eps_l (already times i for
 *                  super-operators, mu.py:130-132); zero matrix if pulse l does
 *                  not drive objective k.
 *   term2pulse   : [K][M] int32: -1 = drift (coefficient 1), l>=0 = pulse l,
 *                  -2 = padding (coefficient 0).
 *   op_norm      : [K][M] upper bound of the 1-norm of each term (used to pick
 *                  the Taylor degree / scaling of exp(A dt) v).
 *   dt           : [NT]  tlist[n+1]-tlist[n]
 *   shape        : [L][NT] update shape S_l on the intervals, in [0,1]
 *   lambda_a     : [L]
 *   pulses       : [L][NT] float64
 *   state stores : [NT+1][K][N] complex128 (time-major)
 */
#ifndef KROTOV_B200_H
#define KROTOV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } kq_c128;

typedef enum {
  KQ_OK = 0,
  KQ_ERR_ARG = -1,        /* invalid argument / unsupported size */
  KQ_ERR_CUDA = -2,       /* CUDA runtime error (message has details) */
  KQ_ERR_UNSUPPORTED = -3,/* configuration outside the built kernels */
  KQ_ERR_EXCHANGE = -4    /* cross-CTA / cross-GPU exchange timed out */
} kq_status;

/* Sparse (CSR) form of the generator terms, their adjoints and the mu matrices, for large
 * state vectors (64 < N <= 1024: Liouville space of notebook 06, N = 625).  Matrix i has
 * rows row_ptr[i*(N+1) .. i*(N+1)+N] (offsets relative to mat_off[i]) into col / val.
 * Numbering: ops[k][m] -> k*M + m;  ops_adj[k][m] -> K*M + k*M + m;
 * mu[k][l] -> 2*K*M + k*L + l.  All arrays are DEVICE pointers. */
typedef struct kq_sparse {
  const int32_t* row_ptr;
  const int64_t* mat_off;
  const int32_t* col;
  const kq_c128* val;
  /* optional dictionary-coded copy (NULL / 0 if absent): val[p] = dict[code16[p]],
   * col16[p] = col[p]; n_dict <= 65536.  Liouvillians repeat few distinct values, so a
   * CTA can keep all matrices of its objective in shared memory (4 bytes per non-zero). */
  const uint16_t* col16;
  const uint16_t* code16;
  const kq_c128* dict;
  int32_t n_dict;
  /* HOST-side hints: the largest number of non-zeros one CTA would stage, over the
   * objectives: generator terms + mu (update sweep) / generator or adjoint terms
   * (propagation sweeps).  0 = do not stage. */
  int32_t stage_nnz_update;
  int32_t stage_nnz_prop;
} kq_sparse;

typedef struct {
  int32_t K;          /* objectives held by this rank */
  int32_t N;          /* state length */
  int32_t NT;         /* number of time intervals (nt - 1) */
  int32_t L;          /* number of pulses */
  int32_t M;          /* generator terms per objective (padded) */
  int32_t is_super;   /* 0: d/dt psi = -i H psi ; 1: d/dt rho = L rho */
  const kq_c128* ops;
  const kq_c128* ops_adj;
  const kq_c128* mu;
  const int32_t* term2pulse;
  const double*  op_norm;
  const double*  dt;
  const double*  shape;
  const double*  lambda_a;
  int32_t real_ops;   /* 1: every matrix in ops/ops_adj AND mu has zero imaginary
                         part (real Hamiltonians): kernels use the purely imaginary
                         form of f*A in Hilbert space; 0: general complex */
  int32_t reserved;   /* kq_sweep_forward_update: number of time windows of the
                         time-parallel update sweep (0 = as few as fit shared memory) */
  int32_t update_sweep; /* which fast update sweep to prefer where both apply (N = 3, 4, few
                         objectives): 0 = library default (time-parallel fixed point),
                         1 = delta-polynomial sequential sweep (strongly coupled problems:
                         the fixed point needs many rounds there), 2 = fixed point.
                         N > 4 always uses the delta-polynomial sweep if it fits. */
  int32_t row_nnz;    /* 0 = unknown (dense rows), else the largest number of non-zero
                         columns in a row of the union pattern of one objective's terms,
                         the diagonal included: rows with <= 4 entries (Lambda systems,
                         transmon ladders) select the entries-in-registers update sweep */
  const struct kq_sparse* sparse;  /* NULL, or the CSR form of all matrices: required for
                         N > 64; for smaller N it selects the row-per-thread CSR kernels if
                         ops is NULL (sparse generators beyond the delta-polynomial family,
                         e.g. a 17-level transmon) and is ignored otherwise */
} kq_problem;

/* Cross-GPU exchange descriptor for the per-time-step reduction of the pulse
 * update (optimize.py:454-470 sums over ALL objectives).  NULL = single GPU.
 * slots[r] is rank r's exchange buffer (kq_comm_slot_bytes() bytes, zeroed
 * once) mapped into this process (CUDA IPC / symmetric memory); every rank
 * writes its partial sums into every peer's buffer and reads only its own.
 * All ranks must call kq_sweep_forward_update with the same `epoch`. */
typedef struct {
  int32_t rank;
  int32_t world;
  void* const* slots;   /* DEVICE array of `world` device pointers */
} kq_comm;

/* chi_constructor kinds lowered to the device (functionals.py). */
enum { KQ_CHI_RE = 0, KQ_CHI_SS = 1, KQ_CHI_SM = 2, KQ_CHI_HS = 3 };

int kq_version(void);
const char* kq_last_error(void);

/* Library options (process-wide).  "picard" (default 1): the fused sweep of
 * kq_sweep_forward_update uses the time-parallel fixed-point kernels where the
 * problem allows (0: always the sequential kernels); "picard_maxit" (default
 * 64): rounds before the time-parallel sweep gives up; "picard_timing": per
 * phase cycle counts in workspace status words 16..25; "picard_history"
 * (default 1): kq_krotov_iteration starts the fixed-point iteration from the
 * extrapolation of the updates of the calls before (kept in the workspace).
 * "cooperative_launch" (default 1) / "programmatic_launch" (default 0,
 * experimental): with 0 / 1 the kernels that exchange data between CTAs are
 * launched with cudaLaunchKernelEx + programmatic stream serialization instead
 * of cudaLaunchCooperativeKernel (co-residency is still checked against the
 * occupancy limit; the kernels wait for the preceding launches before they
 * touch memory).
 * "lanes" (default 1): the entries-in-registers kernels (csrc/kq_lanes.cuh) for few
 * objectives with several controls or sparse rows, 0: the generic kernels instead;
 * "sat" (default 1): the many-objective update sweep (csrc/kq_sat.cuh: N = 2, real
 * generator, more objectives than one CTA of the sequential kernel holds), 0: the
 * sequential kernel with its per-step slot exchange.
 * "dpoly" (default 1): the delta-polynomial Krotov iteration (csrc/kq_dpoly.cuh: few
 * objectives, one control, first order; per time step ONE small matrix-vector product
 * with a step propagator that is a polynomial in the deviation of the pulse from an
 * anchor pulse; the polynomials are kept in the workspace and rebuilt only when the
 * pulse has drifted out of their radius; the backward sweep is a time-parallel product
 * of the same polynomials) is used where kq_problem.update_sweep / the state dimension
 * ask for it; 0: never; 2: wherever it fits.  kq_sweep_forward_update queues the
 * sequential Taylor kernels behind it as an in-stream conditional fall-back;
 * kq_krotov_iteration reports a declined iteration through diag_out / the status words
 * (the caller repeats it with the sweep calls).
 * "time_parallel" (default 1): propagation
 * sweeps under known pulses (kq_propagate_forward, kq_sweep_backward*) are cut
 * into time segments that run concurrently (segment propagators -> boundary
 * states -> states); 0 selects the purely sequential sweep. */
int kq_set_option(const char* name, int value);

/* Cross-GPU exchange buffers (one process per GPU).  kq_comm_alloc allocates
 * `bytes` of zeroed device memory on the current device and returns a 64-byte
 * CUDA IPC handle to pass to the peer processes (e.g. through
 * torch.distributed.all_gather_object); kq_comm_open maps a peer's buffer.
 * Every rank allocates kq_comm_slot_bytes() this way; kq_comm.slots[r] is
 * rank r's buffer as mapped into the calling process. */
int kq_comm_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int kq_comm_open(const unsigned char* handle64, void** ptr);
int kq_comm_close(void* ptr);
int kq_comm_free(void* ptr);

/* Stream-ordered barrier over the ranks of `comm` (one tiny kernel: flag-tagged
 * stores into every peer's exchange buffer, bounded spin on the local one).
 * `tag` must be non-zero and increase with every call; `workspace` (may be
 * NULL) receives KQ_ERR_EXCHANGE in its status word on time-out. */
int kq_comm_barrier(const kq_comm* comm, uint32_t tag, void* workspace,
                    void* stream);

/* Bytes of zero-initialised device workspace kq_sweep_forward_update needs
 * (status word + cross-CTA exchange slots). */
size_t kq_workspace_bytes(const kq_problem* p);
/* Diagnostics: byte offset in the workspace of the header (128 bytes reserved) of the
 * delta-polynomial iteration {int32 J, m, rebuild, usable, anchor_epoch, valid_epoch,
 * builds, reuses; double radius, last_max, anchor_max, dtmax, o0, o1}. */
size_t kq_dpoly_header_offset(const kq_problem* p);
size_t kq_comm_slot_bytes(const kq_problem* p);

/* Forward propagation over the whole grid under `pulses`:
 * store[0] = state0, store[n+1] = exp(f A_n dt_n) store[n]; stateT = final.
 * `store` and `stateT` may be NULL (not both). */
int kq_propagate_forward(const kq_problem* p, const double* pulses,
                         const kq_c128* state0, kq_c128* stateT,
                         kq_c128* store, void* stream);

/* Backward propagation of chiT (already normalised) under the adjoint
 * generator: X[NT] = chiT, X[n] = exp(conj(f) A^dag_n dt_n) X[n+1]. */
int kq_sweep_backward(const kq_problem* p, const double* guess_pulses,
                      const kq_c128* chiT, kq_c128* X, void* stream);

/* 'gather' multi-GPU mode: backward propagation of the objectives
 * [k_lo, k_lo+k_cnt) only into this rank's store X_peers[self], followed by a
 * copy kernel that writes that block of columns into the [NT+1][K][N] stores
 * of all peers (X_peers: HOST array of n_peer device pointers mapped with
 * kq_comm_alloc/kq_comm_open; wide P2P stores over NVLink).  After a
 * kq_comm_barrier every GPU holds the complete X and runs the fused sweep
 * replicated, without any per-step exchange.  chiT is the full [K][N] array. */
int kq_sweep_backward_range(const kq_problem* p, const double* guess_pulses,
                            const kq_c128* chiT, void* const* X_peers,
                            int32_t n_peer, int32_t self, int32_t k_lo,
                            int32_t k_cnt, void* stream);

/* Fused sequential sweep: for n = 0..NT-1
 *   d_l   = Im sum_k [ chi_norms[k] <X[n][k]| mu_lk |phi_k>
 *                      + 0.5 sigma[n] <dphi_k| mu_lk |phi_k> ]
 *   opt_l[n] = guess_l[n] + (S_l[n]/lambda_l) d_l ;  g_a[l] += (S/lambda) d_l^2 dt
 *   phi_k <- exp(f A_k(opt[:,n]) dt_n) phi_k ; dphi_k = phi_k - Phi0[n+1][k]
 * sigma/Phi0/Phi1 NULL = first order.  Phi1 (if given) receives all forward
 * states.  g_a is overwritten.  `workspace` = kq_workspace_bytes() zeroed
 * bytes, `epoch` must increase by one with every call sharing a workspace. */
int kq_sweep_forward_update(const kq_problem* p, const double* guess_pulses,
                            double* opt_pulses, const kq_c128* X,
                            const double* chi_norms, const kq_c128* phi0,
                            kq_c128* phiT, const double* sigma,
                            const kq_c128* Phi0, kq_c128* Phi1, double* g_a,
                            const kq_comm* comm, void* workspace,
                            uint32_t epoch, void* stream);

/* One complete Krotov iteration (optimize.py:393-508) in ONE launch of the
 * time-parallel kernel family (csrc/kq_picard.cuh): boundary condition
 * (chi_kind = KQ_CHI_* evaluated from targets/weights/tau_in/phiT_in, or -1 with
 * the normalised chiT [K][N] + chi_norms [K] of a host chi_constructor),
 * backward sweep under guess_pulses (X receives all backward states if not
 * NULL), pulse update + forward sweep (the sequential chain is replaced by a
 * causal fixed-point iteration that is parallel in time), tau_out[k] =
 * <target_k|phi_k(T)> (if targets and tau_out are given), phiT_out, g_a.
 * chi_out / chi_norms_out (may be NULL) receive the normalised boundary states.
 * prev_guess_pulses (may be NULL, may alias opt_pulses) = the guess pulses of
 * the Krotov iteration before: the fixed-point iteration then starts from
 * guess + (guess - prev_guess) instead of guess (a hint only: the result does
 * not depend on it beyond rounding).  The kernel also keeps the last three
 * updates (opt - guess) in `workspace`; when guess_pulses is the buffer the
 * previous call wrote its opt_pulses to, the first iterate is the guess plus
 * the polynomial extrapolation of those updates instead (about one to two
 * rounds fewer; option "picard_history", default 1).
 * sigma/Phi0/Phi1 as in kq_sweep_forward_update.  tau_in/phiT_in must not alias
 * tau_out/phiT_out.
 * Objectives sharded over GPUs (comm != NULL, comm->world > 1; one process per
 * GPU, every rank calls with the same epoch): this rank holds K <=
 * ceil(K_total / world) objectives; all ranks run the same launch geometry.  In
 * every fixed-point round the CTA that owns a time slice adds the sums of the
 * other GPUs to its own -- one 16-byte flag-tagged store per value into every
 * peer's exchange buffer (comm->slots[r], kq_comm_slot_bytes() bytes) over
 * NVLink, polling of the local buffer only, summation in rank order -- so the
 * reduction over ALL objectives of optimize.py:454-470 crosses the GPUs once per
 * round (nt doubles) instead of once per time step, and every rank obtains
 * bit-identical pulses.  No NCCL call or host synchronisation inside the
 * iteration.  tau_out / phiT_out / X hold this rank's objectives.  For
 * KQ_CHI_SM tau_sum points to sum_j w_j tau_j over ALL ranks (one complex value
 * on the device, e.g. from an NCCL all-reduce on the same stream); NULL otherwise.
 * Two more kernel families stand behind the same call (several launches, one call):
 * the delta-polynomial iteration ("dpoly" option: few objectives, one control,
 * N <= 16) and the entries-in-registers chains (csrc/kq_lanes.cuh, "lanes" option:
 * few objectives, up to four controls and five terms, N <= 32 with at most four
 * non-zero entries per row -- kq_problem.row_nnz --, first order, one GPU: chi
 * boundary, time-parallel backward sweep, pre-pass, update chain, tau).
 * Returns KQ_ERR_UNSUPPORTED when the problem is outside these
 * families (fixed point: N in 2..4, M = 2, L = 1, state stores fit shared memory):
 * use the four-call sequence then.  If the fixed-point
 * iteration does not converge ("picard_maxit" option) the outputs are left
 * untouched, workspace status word 1 is set to `epoch` and word 3 to the first
 * such epoch; word 2 holds the number of fixed-point rounds of the last
 * converged call.  diag_out (may be NULL): 4 int32 written by the kernel,
 * {exchange status, `epoch` if not converged else 0, rounds, 0}, so that a
 * caller can fetch results and status with one device->host copy. */
int kq_krotov_iteration(const kq_problem* p, int chi_kind, int32_t K_total,
                        const kq_c128* targets, const double* weights,
                        const kq_c128* chiT, const double* chi_norms,
                        const kq_c128* tau_in, const kq_c128* phiT_in,
                        const double* guess_pulses,
                        const double* prev_guess_pulses, double* opt_pulses,
                        const kq_c128* phi0, kq_c128* phiT_out,
                        kq_c128* tau_out, kq_c128* X, kq_c128* chi_out,
                        double* chi_norms_out, const double* sigma,
                        const kq_c128* Phi0, kq_c128* Phi1, double* g_a,
                        int32_t* diag_out, const kq_comm* comm,
                        const kq_c128* tau_sum, void* workspace, uint32_t epoch,
                        void* stream);

/* Boundary condition chi_k(T) for the built-in functionals, followed by the
 * normalisation of optimize.py:407-410 (L2 / Frobenius norm):
 * chi_out[k] = chi_k/||chi_k||, chi_norms[k] = ||chi_k||.
 * weights may be NULL (all 1).  K_total = number of objectives over all
 * ranks (the 1/N prefactors of functionals.py use the global count);
 * tau_sum = sum_j w_j tau_j over ALL objectives for KQ_CHI_SM (device pointer
 * to one complex value), ignored otherwise. */
int kq_chi_boundary(const kq_problem* p, int kind, int32_t K_total,
                    const kq_c128* phiT, const kq_c128* targets,
                    const kq_c128* tau, const double* weights,
                    const kq_c128* tau_sum, kq_c128* chi_out,
                    double* chi_norms, void* stream);

/* Asynchronous device->host copy of one iteration's packed results (pulses |
 * g_a | tau | status words, written by kq_krotov_iteration into caller-owned
 * device memory) into pinned host memory, on `stream` -- normally a copy
 * stream that waits on an event recorded behind that iteration's launch, so
 * that iterations already queued behind it do not delay the hooks
 * (reference: the per-iteration results handed to info_hook,
 * optimize.py:511-533). */
int kq_fetch_results(void* dst_host, const void* src_device, size_t nbytes,
                     void* stream);

/* out[k] = <a_k | b_k> for K vectors of length N (tau_vals). */
int kq_overlaps(int32_t K, int32_t N, const kq_c128* a, const kq_c128* b,
                kq_c128* out, void* stream);

/* Introspection for tests / bench: kernel family the sequential update sweep of
 * kq_sweep_forward_update uses for a problem on a 148-SM device (0 = thread per
 * objective, 1 = lane per row, 2 = row per thread on CSR matrices, 3 = entries in
 * registers -- csrc/kq_lanes.cuh --, 4 = many-objective sweep -- csrc/kq_sat.cuh)
 * and its launch geometry. */
int kq_plan(const kq_problem* p, int32_t* family, int32_t* grid,
            int32_t* block, int32_t* smem_bytes);

/* Launch geometry of kq_krotov_iteration on a 148-SM device: CTAs, threads per
 * CTA, time steps per thread (chunk) and dynamic shared memory; returns
 * KQ_ERR_UNSUPPORTED if the problem is outside that kernel family.  Host-only
 * (no device needed). */
int kq_plan_fused(const kq_problem* p, int32_t* grid, int32_t* block,
                  int32_t* chunk, int32_t* smem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* KROTOV_B200_H */
