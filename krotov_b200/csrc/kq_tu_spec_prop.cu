// Specialised propagation sweeps (kq_spec.cuh): instantiations with complex
// generators, including the time-parallel (segmented) variant.
#include "kq_tu_spec_prop.inc"
KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_spec_prop)
int kq_launch_prop_spec(const KqSweepArgs& a, const KqPlan& pl, int fsel, int nseg,
                        cudaStream_t st) {
  return dispatch_prop<cplx>(a, pl, fsel, nseg, st);
}
