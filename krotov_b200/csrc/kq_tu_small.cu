// Generic thread-per-objective kernels (kq_small.cuh): instantiations.
#include "kq_host.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_small)

namespace {
template <int N>
int launch_prop_small(const KqSweepArgs& a, const Plan& pl, int fsel, cudaStream_t st) {
  void* params[] = {(void*)&a};
  switch (fsel) {
    case 0: return launch(k_prop_small<N, 0>, pl, false, st, params);
    case 1: return launch(k_prop_small<N, 1>, pl, false, st, params);
    default: return launch(k_prop_small<N, 2>, pl, false, st, params);
  }
}

}  // namespace

int kq_launch_prop_small(const KqSweepArgs& a, const KqPlan& pl, int fsel, cudaStream_t st) {
  switch (a.N) {
    case 1: return launch_prop_small<1>(a, pl, fsel, st);
    case 2: return launch_prop_small<2>(a, pl, fsel, st);
    case 3: return launch_prop_small<3>(a, pl, fsel, st);
    default: return launch_prop_small<4>(a, pl, fsel, st);
  }
}

namespace {
template <int N>
int launch_fwupd_small(const KqSweepArgs& a, const Plan& pl, int fsel, bool second,
                       cudaStream_t st) {
  void* params[] = {(void*)&a};
  const bool coop = pl.grid > 1;
  if (N <= 2 && pl.block > 256) {
    constexpr int BT = (N <= 2) ? 1024 : 256;
    if (fsel == 0) {
      return second ? launch(k_fwupd_small<N, 0, true, BT>, pl, coop, st, params)
                    : launch(k_fwupd_small<N, 0, false, BT>, pl, coop, st, params);
    }
    return second ? launch(k_fwupd_small<N, 2, true, BT>, pl, coop, st, params)
                  : launch(k_fwupd_small<N, 2, false, BT>, pl, coop, st, params);
  }
  if (fsel == 0) {
    return second ? launch(k_fwupd_small<N, 0, true, 256>, pl, coop, st, params)
                  : launch(k_fwupd_small<N, 0, false, 256>, pl, coop, st, params);
  }
  return second ? launch(k_fwupd_small<N, 2, true, 256>, pl, coop, st, params)
                : launch(k_fwupd_small<N, 2, false, 256>, pl, coop, st, params);
}

}  // namespace

int kq_launch_fwupd_small(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                          cudaStream_t st) {
  switch (a.N) {
    case 1: return launch_fwupd_small<1>(a, pl, fsel, second, st);
    case 2: return launch_fwupd_small<2>(a, pl, fsel, second, st);
    case 3: return launch_fwupd_small<3>(a, pl, fsel, second, st);
    default: return launch_fwupd_small<4>(a, pl, fsel, second, st);
  }
}
