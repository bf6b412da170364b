// Lane = (objective, row) update sweep for few objectives with small or sparse generators
// (kq_lanes.cuh): launches.
#include "kq_host.cuh"
#include "kq_lanes.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_lanes)

namespace {
template <int FSEL, bool MULTI>
int launch_lanes(const KqSweepArgs& a, const KqLanes& d, cudaStream_t st) {
  const int bt = d.W * 32;
  k_rows_prep<FSEL><<<a.NT, bt, 0, st>>>(a, d);
  KQ_CUDA(cudaGetLastError());
  switch (d.NZ) {
    case 2: k_fwupd_rows<2, FSEL, MULTI><<<1, bt, 0, st>>>(a, d); break;
    case 3: k_fwupd_rows<3, FSEL, MULTI><<<1, bt, 0, st>>>(a, d); break;
    default: k_fwupd_rows<4, FSEL, MULTI><<<1, bt, 0, st>>>(a, d); break;
  }
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}
}  // namespace

int kq_launch_lanes(const KqSweepArgs& a, const KqLanes& d, int fsel, cudaStream_t st) {
  if (d.W > 1) return fsel == 0 ? launch_lanes<0, true>(a, d, st) : launch_lanes<2, true>(a, d, st);
  return fsel == 0 ? launch_lanes<0, false>(a, d, st) : launch_lanes<2, false>(a, d, st);
}

namespace {
template <int FSEL>
int launch_prop_rows(const KqSweepArgs& a, const KqLanes& d, int grid, cudaStream_t st) {
  const int bt = d.W * 32;
  switch (d.NZ) {
    case 2: k_prop_rows<2, FSEL><<<grid, bt, 0, st>>>(a, d); break;
    case 3: k_prop_rows<3, FSEL><<<grid, bt, 0, st>>>(a, d); break;
    default: k_prop_rows<4, FSEL><<<grid, bt, 0, st>>>(a, d); break;
  }
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}
}  // namespace

// propagation sweep: `n_task` tasks of d.NP lanes, d.W warps per CTA
int kq_launch_prop_lanes(const KqSweepArgs& a, const KqLanes& d, int fsel, int n_task,
                         cudaStream_t st) {
  const int per_cta = d.W * d.G;
  const int grid = (n_task + per_cta - 1) / per_cta;
  switch (fsel) {
    case 0: return launch_prop_rows<0>(a, d, grid, st);
    case 1: return launch_prop_rows<1>(a, d, grid, st);
    default: return launch_prop_rows<2>(a, d, grid, st);
  }
}
