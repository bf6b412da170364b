// Argument block of the lane = (objective, row) update sweep (kq_lanes.cuh), shared with kq_abi.cu.
#pragma once
#include "kq_common.cuh"

#define KQ_LN_LMAX 4    // controls at most
#define KQ_LN_NZMAX 4   // non-zero entries per row at most (union over the terms, plus the diagonal)
#define KQ_LN_WMAX 8    // warps at most
#define KQ_LN_SC 10     // doubles per scalar record: S/lambda [4] | guess [4] | dt | 0

struct KqLanes {
  cplx* zeta;     // [NT][L + 1][W * 32]: zeta_l of every lane | exp(c0 dt)
  double* scal;   // [NT][KQ_LN_SC]
  int NP;         // lanes per objective (N rounded up to a power of two)
  int G, W;       // objectives per warp, warps
  int NZ;         // entries kept per row (template variant)
  int span;       // lanes the in-warp butterfly covers (power of two)
};
