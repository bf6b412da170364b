// Specialised propagation sweeps (kq_spec.cuh): instantiations, including
// the time-parallel (segmented) variant.
#include "kq_host.cuh"
#include "kq_spec.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_spec_prop)

namespace {
template <int N, int NVEC>
int go(const KqSweepArgs& a, const KqPlan& pl, int fsel, cudaStream_t st) {
  void* params[] = {(void*)&a};
  switch (fsel) {
    case 0: return launch(k_prop_spec<N, 0, NVEC>, pl, false, st, params);
    case 1: return launch(k_prop_spec<N, 1, NVEC>, pl, false, st, params);
    default: return launch(k_prop_spec<N, 2, NVEC>, pl, false, st, params);
  }
}

template <int N>
int segmented(KqSweepArgs a, KqPlan pl, int fsel, int nseg, cudaStream_t st) {
  pl.grid_y = nseg;
  a.seg_pass = 1;
  int rc = go<N, N>(a, pl, fsel, st);
  if (rc) return rc;
  k_seg_chain<N><<<(a.k_cnt + 127) / 128, 128, 0, st>>>(a, nseg);
  KQ_CUDA(cudaGetLastError());
  a.seg_pass = 2;
  return go<N, 1>(a, pl, fsel, st);
}
}  // namespace

int kq_launch_prop_spec(const KqSweepArgs& a, const KqPlan& pl, int fsel, int nseg,
                        cudaStream_t st) {
  if (nseg > 1) {
    switch (a.N) {
      case 2: return segmented<2>(a, pl, fsel, nseg, st);
      case 3: return segmented<3>(a, pl, fsel, nseg, st);
      default: return segmented<4>(a, pl, fsel, nseg, st);
    }
  }
  switch (a.N) {
    case 2: return go<2, 1>(a, pl, fsel, st);
    case 3: return go<3, 1>(a, pl, fsel, st);
    default: return go<4, 1>(a, pl, fsel, st);
  }
}
