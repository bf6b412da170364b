// Launch geometry of the lane-per-row kernels (kq_warp.cuh).
#pragma once
struct KqWarpGeom {
  int R;              // lanes per objective
  int G;              // objectives per warp
  int terms_in_smem;  // generator terms resident in shared memory
  int mu_in_smem;
  int obj_stride;     // per-objective shared memory, in cplx units
};
