// Time-parallel fused update/forward sweep (kq_picard.cuh), N = 3: instantiations.
#include "kq_host.cuh"
#include "kq_picard.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_picard3)

int kq_launch_fwupd_picard3(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                            bool real, cudaStream_t st) {
  void* params[] = {(void*)&a};
  if (real && fsel == 0) {
    return second ? launch(k_fwupd_picard<3, 0, true, double>, pl, true, st, params)
                  : launch(k_fwupd_picard<3, 0, false, double>, pl, true, st, params);
  }
  if (fsel == 0) {
    return second ? launch(k_fwupd_picard<3, 0, true, cplx>, pl, true, st, params)
                  : launch(k_fwupd_picard<3, 0, false, cplx>, pl, true, st, params);
  }
  return second ? launch(k_fwupd_picard<3, 2, true, cplx>, pl, true, st, params)
                : launch(k_fwupd_picard<3, 2, false, cplx>, pl, true, st, params);
}
