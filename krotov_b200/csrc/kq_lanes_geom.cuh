// Argument block of the one-warp update sweep (kq_lanes.cuh), shared with kq_abi.cu.
#pragma once
#include "kq_common.cuh"

#define KQ_LN_LMAX 4    // controls at most
#define KQ_LN_SC 10     // doubles per scalar record: S/lambda [4] | guess [4] | dt | 0

struct KqLanes {
  cplx* zeta;     // [NT][L][32]
  double* scal;   // [NT][KQ_LN_SC]
  int NP;         // lanes per objective (N rounded up to a power of two)
  int span;       // K * NP rounded up to a power of two (<= 32)
};
