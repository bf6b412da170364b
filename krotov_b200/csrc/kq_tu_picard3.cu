// Time-parallel Krotov iteration (kq_picard.cuh), N = 3: instantiations.
#include "kq_host.cuh"
#include "kq_picard.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_picard3)

namespace {
template <int FSEL, bool SECOND, typename G>
int by_chunk(const KqSweepArgs& a, const KqPlan& pl, cudaStream_t st) {
  void* params[] = {(void*)&a};
  // run-time chunk loops: for N >= 3 the unrolled variants (registers, code size)
  // are no faster (N = 3) or much slower (N = 4: 2.1x) than the rolled loops
  return launch(k_krotov_picard<3, FSEL, SECOND, G, 0>, pl, true, st, params);
}
}  // namespace

int kq_launch_picard3(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool real,
                      cudaStream_t st) {
  if (real && fsel == 0)
    return second ? by_chunk<0, true, double>(a, pl, st) : by_chunk<0, false, double>(a, pl, st);
  if (fsel == 0)
    return second ? by_chunk<0, true, cplx>(a, pl, st) : by_chunk<0, false, cplx>(a, pl, st);
  return second ? by_chunk<2, true, cplx>(a, pl, st) : by_chunk<2, false, cplx>(a, pl, st);
}
