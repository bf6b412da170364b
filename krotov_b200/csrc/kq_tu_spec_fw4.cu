// Specialised fused update/forward sweep, N = 4, generator element type cplx
// (kq_spec.cuh): instantiations.
#include "kq_host.cuh"
#include "kq_spec.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_spec_fw4)

int kq_launch_fwupd_spec4(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                          cudaStream_t st) {
  void* params[] = {(void*)&a};
  const bool coop = pl.grid > 1;
  if (fsel == 0) {
    return second ? launch(k_fwupd_spec<4, 0, true, 256, cplx>, pl, coop, st, params)
                  : launch(k_fwupd_spec<4, 0, false, 256, cplx>, pl, coop, st, params);
  }
  return second ? launch(k_fwupd_spec<4, 2, true, 256, cplx>, pl, coop, st, params)
                : launch(k_fwupd_spec<4, 2, false, 256, cplx>, pl, coop, st, params);
}
