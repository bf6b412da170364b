// Geometry and argument blocks of the delta-polynomial update sweep (kq_dpoly.cuh), shared
// by the kernels and the host-side planning in kq_abi.cu.
#pragma once
#include "kq_common.cuh"

#define KQ_DP_JMAX 8      // highest polynomial degree in delta
#define KQ_DP_CMAX 4      // columns per lane at most
#define KQ_DP_MAXLANES 512
#define KQ_DP_RINGMAX 16   // stages of the sweep kernel's shared-memory ring at most

struct KqDpHeader {       // in the caller's workspace, survives between calls
  int J;                  // degree used by the call in progress
  int valid_epoch;        // epoch whose largest update is in last_max
  int m;                  // Taylor degree of the step propagators (largest step)
  int pad;
  double delta_bound;     // bound on |delta| the call in progress was built for
  double last_max;        // largest |opt - guess| of epoch valid_epoch
};

struct KqDpoly {
  cplx* rec;              // [NT][rec_stride] step records
  int rec_stride;         // complex numbers per record (sized for KQ_DP_JMAX)
  int NL, Q, C, Npad;     // lanes = K (N+1) Q, lanes per row, columns per lane, padded row length
  int TPC;                // build kernel: time steps per CTA
  int ring;               // sweep kernel: capacity of the shared-memory ring (complex numbers)
  KqDpHeader* hdr;
};

