// Time-parallel Krotov iteration (kq_picard.cuh), N = 2: instantiations.
#include "kq_host.cuh"
#include "kq_picard.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_picard2)

namespace {
template <int FSEL, bool SECOND, typename G>
int by_chunk(const KqSweepArgs& a, const KqPlan& pl, cudaStream_t st) {
  void* params[] = {(void*)&a};
  // N = 2: compile-time chunk lengths (unrolled, step operators kept in registers)
  // halve pass A / pass B against the run-time loop
  switch (a.pic_W) {
    case 2: return launch(k_krotov_picard<2, FSEL, SECOND, G, 2>, pl, true, st, params);
    case 4: return launch(k_krotov_picard<2, FSEL, SECOND, G, 4>, pl, true, st, params);
    case 8: return launch(k_krotov_picard<2, FSEL, SECOND, G, 8>, pl, true, st, params);
    default: return launch(k_krotov_picard<2, FSEL, SECOND, G, 0>, pl, true, st, params);
  }
}
}  // namespace

int kq_launch_picard2(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool real,
                      cudaStream_t st) {
  if (real && fsel == 0)
    return second ? by_chunk<0, true, double>(a, pl, st) : by_chunk<0, false, double>(a, pl, st);
  if (fsel == 0)
    return second ? by_chunk<0, true, cplx>(a, pl, st) : by_chunk<0, false, cplx>(a, pl, st);
  return second ? by_chunk<2, true, cplx>(a, pl, st) : by_chunk<2, false, cplx>(a, pl, st);
}
