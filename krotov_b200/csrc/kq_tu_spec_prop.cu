// Specialised propagation sweeps (kq_spec.cuh): instantiations.
#include "kq_host.cuh"
#include "kq_spec.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_spec_prop)

namespace {
template <int N>
int go(const KqSweepArgs& a, const KqPlan& pl, int fsel, cudaStream_t st) {
  void* params[] = {(void*)&a};
  switch (fsel) {
    case 0: return launch(k_prop_spec<N, 0>, pl, false, st, params);
    case 1: return launch(k_prop_spec<N, 1>, pl, false, st, params);
    default: return launch(k_prop_spec<N, 2>, pl, false, st, params);
  }
}
}  // namespace

int kq_launch_prop_spec(const KqSweepArgs& a, const KqPlan& pl, int fsel, cudaStream_t st) {
  switch (a.N) {
    case 2: return go<2>(a, pl, fsel, st);
    case 3: return go<3>(a, pl, fsel, st);
    default: return go<4>(a, pl, fsel, st);
  }
}
