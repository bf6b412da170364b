// Row-per-thread CSR kernels for 64 < N <= 1024 (kq_csr.cuh): launches.
#include "kq_host.cuh"
#include "kq_csr.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_csr)

namespace {
template <bool STAGED>
int launch_csr(const KqSweepArgs& a, const KqCsr& s, const KqPlan& pl, int fsel, bool update,
               cudaStream_t st) {
  void* params[] = {(void*)&a, (void*)&s};
  if (update) {
    // the objectives' CTAs exchange partial sums: they must be co-resident
    const bool coop = pl.grid > 1;
    switch (fsel) {
      case 0: return launch(k_sweep_csr<0, true, STAGED>, pl, coop, st, params);
      default: return launch(k_sweep_csr<2, true, STAGED>, pl, coop, st, params);
    }
  }
  switch (fsel) {
    case 0: return launch(k_sweep_csr<0, false, STAGED>, pl, false, st, params);
    case 1: return launch(k_sweep_csr<1, false, STAGED>, pl, false, st, params);
    default: return launch(k_sweep_csr<2, false, STAGED>, pl, false, st, params);
  }
}
}  // namespace

int kq_launch_csr(const KqSweepArgs& a, const KqCsr& s, const KqPlan& pl, int fsel, bool update,
                  bool staged, cudaStream_t st) {
  return staged ? launch_csr<true>(a, s, pl, fsel, update, st)
                : launch_csr<false>(a, s, pl, fsel, update, st);
}
