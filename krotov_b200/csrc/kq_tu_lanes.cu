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
