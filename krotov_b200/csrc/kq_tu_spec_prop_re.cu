// Specialised propagation sweeps for real generator matrices in Hilbert
// space (purely imaginary f*A: half the multiply work), kq_spec.cuh.
#include "kq_tu_spec_prop.inc"
KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_spec_prop_re)
int kq_launch_prop_spec_re(const KqSweepArgs& a, const KqPlan& pl, int fsel, int nseg,
                           cudaStream_t st) {
  return dispatch_prop<double>(a, pl, fsel, nseg, st);
}
