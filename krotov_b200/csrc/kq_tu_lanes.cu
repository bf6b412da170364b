// One-warp update sweep for few objectives with several controls (kq_lanes.cuh): launches.
#include "kq_host.cuh"
#include "kq_lanes.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_lanes)

namespace {
template <int N>
int launch_lanes(const KqSweepArgs& a, const KqLanes& d, int fsel, cudaStream_t st) {
  if (fsel == 0)
    k_fwupd_lanes<N, 0><<<1, 32, 0, st>>>(a, d);
  else
    k_fwupd_lanes<N, 2><<<1, 32, 0, st>>>(a, d);
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}
}  // namespace

int kq_launch_lanes(const KqSweepArgs& a, const KqLanes& d, int fsel, cudaStream_t st) {
  k_lanes_prep<<<(a.NT + 7) / 8, 256, 0, st>>>(a, d);
  KQ_CUDA(cudaGetLastError());
  switch (a.N) {
    case 2: return launch_lanes<2>(a, d, fsel, st);
    case 3: return launch_lanes<3>(a, d, fsel, st);
    default: return launch_lanes<4>(a, d, fsel, st);
  }
}
