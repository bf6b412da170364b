// Lane-per-row kernels (kq_warp.cuh): instantiations for mode 0.
#include "kq_host.cuh"
#include "kq_warp.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_warp0)

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

int kq_launch_warp0(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool update,
                    cudaStream_t st) {
  if (pl.rpl == 1) return launch_warp_mode<1, 0>(a, pl, fsel, second, update, st);
  const bool allsm = pl.geom.terms_in_smem && (!update || pl.geom.mu_in_smem);
  return allsm ? launch_warp_mode<2, 1>(a, pl, fsel, second, update, st)
               : launch_warp_mode<2, 0>(a, pl, fsel, second, update, st);
}
