// Geometry and argument blocks of the delta-polynomial Krotov iteration (kq_dpoly.cuh),
// shared by the kernels and the host-side planning in kq_abi.cu.
#pragma once
#include "kq_common.cuh"

#define KQ_DP_JMAX 8      // highest polynomial degree in Delta
#define KQ_DP_CMAX 4      // columns per lane at most
#define KQ_DP_MAXLANES 512
#define KQ_DP_RINGMAX 16   // stages of the sweep kernel's shared-memory ring at most
#define KQ_DP_SEGMAX 256   // segments of the time-parallel backward sweep at most
#define KQ_DP_PAD 16      // chunk length of the sweep kernel's ring at most; identity records
                          // behind the grid: 2 x this
#define KQ_DP_GROW 6.0     // a rebuild asks for a radius of GROW x the predicted update

// In the caller's workspace; survives between calls (zero = nothing built yet).
struct KqDpHeader {
  int J;                  // degree of the step records
  int m;                  // Taylor degree the records were built with
  int rebuild;            // set by the plan kernel: the build kernel has to run in this call
  int usable;             // set by the plan kernel: the records cover this call
  int anchor_epoch;       // epoch the records were built in (0 = none)
  int valid_epoch;        // epoch whose largest update is in last_max
  int builds, reuses;     // counters (diagnostics)
  double radius;          // the records reproduce U_n(anchor_n + Delta) for |Delta| <= radius
  double last_max;        // largest |opt - guess| of epoch valid_epoch
  double anchor_max;      // max |anchor|
  double dtmax, o0, o1;   // problem constants (max |dt|, operator norm bounds), 0 = not yet known
};

struct KqDpoly {
  cplx* rec;              // [NT + 2 KQ_DP_PAD] step records (E part | zeta part | scalars), stored
                          // compactly for the degree in use
  int rec_stride;         // complex numbers per record at degree KQ_DP_JMAX (allocation)
  int NL, Q, C, Npad;     // lanes = K (N+1) Q, lanes per row, columns per lane, padded row length
  int TPC;                // build kernel: time steps per CTA
  int ring;               // sweep kernel: capacity of the shared-memory ring (complex numbers)
  int nseg, seg_len;      // backward sweep: segments of seg_len steps
  int chain;              // 1: backward states from chi(T) through the records; 0: read from X
  int R2;                 // N rounded up to a power of two (expand kernel: lanes per column)
  int chain_bs;           // expand kernel: segment propagators per shared-memory batch
  int estage_cap;         // segprod / expand: complex numbers per record-staging buffer
  KqDpHeader* hdr;
  double* anchor;         // [NT] pulse the records are expanded around
  cplx* segP;             // [nseg][K][N*N] segment propagators (row-major)
  cplx* X;                // [NT+1][K][N] backward states (chain: written; else: read)
  cplx* chi;              // [K][N] normalised chi(T)
  double* norms;          // [K] ||chi_k(T)||
  cplx* tsum;             // scratch: sum_j w_j tau_j (chis_sm)
};
