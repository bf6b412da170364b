// Specialised fused update/forward sweep, N = 2, generator element type cplx
// (kq_spec.cuh): instantiations.
#include "kq_host.cuh"
#include "kq_spec.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_spec_fw2)

int kq_launch_fwupd_spec2(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                          cudaStream_t st) {
  void* params[] = {(void*)&a};
  const bool coop = pl.grid > 1;
  if (pl.block > 256) {
    if (fsel == 0) {
      return second ? launch(k_fwupd_spec<2, 0, true, 1024, cplx>, pl, coop, st, params)
                    : launch(k_fwupd_spec<2, 0, false, 1024, cplx>, pl, coop, st, params);
    }
    return second ? launch(k_fwupd_spec<2, 2, true, 1024, cplx>, pl, coop, st, params)
                  : launch(k_fwupd_spec<2, 2, false, 1024, cplx>, pl, coop, st, params);
  }
  if (fsel == 0) {
    return second ? launch(k_fwupd_spec<2, 0, true, 256, cplx>, pl, coop, st, params)
                  : launch(k_fwupd_spec<2, 0, false, 256, cplx>, pl, coop, st, params);
  }
  return second ? launch(k_fwupd_spec<2, 2, true, 256, cplx>, pl, coop, st, params)
                : launch(k_fwupd_spec<2, 2, false, 256, cplx>, pl, coop, st, params);
}
