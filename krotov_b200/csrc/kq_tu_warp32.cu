// Lane-per-row kernels (kq_warp.cuh): instantiations for mode 32.
#include "kq_host.cuh"
#include "kq_warp.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_warp32)

namespace {
template <int RPL, int MODE>
int launch_warp_mode(const KqSweepArgs& a, const Plan& pl, int fsel, bool second, bool update,
                     cudaStream_t st) {
  void* params[] = {(void*)&a, (void*)&pl.geom};
  const bool coop = update && pl.grid > 1;
  if (!update) {
    switch (fsel) {
      case 0: return launch(k_sweep_warp<RPL, 0, false, false, MODE>, pl, false, st, params);
      case 1: return launch(k_sweep_warp<RPL, 1, false, false, MODE>, pl, false, st, params);
      default: return launch(k_sweep_warp<RPL, 2, false, false, MODE>, pl, false, st, params);
    }
  }
  if (fsel == 0) {
    return second ? launch(k_sweep_warp<RPL, 0, true, true, MODE>, pl, coop, st, params)
                  : launch(k_sweep_warp<RPL, 0, false, true, MODE>, pl, coop, st, params);
  }
  return second ? launch(k_sweep_warp<RPL, 2, true, true, MODE>, pl, coop, st, params)
                : launch(k_sweep_warp<RPL, 2, false, true, MODE>, pl, coop, st, params);
}

}  // namespace

int kq_launch_warp32(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool update,
                    cudaStream_t st) {
  return launch_warp_mode<1, 32>(a, pl, fsel, second, update, st);
}
