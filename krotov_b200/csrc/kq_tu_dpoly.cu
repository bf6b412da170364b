// Delta-polynomial update sweep (kq_dpoly.cuh): launches.
#include "kq_host.cuh"
#include "kq_dpoly.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_dpoly)

namespace {
template <int NMAX>
int build_and_zeta(const KqSweepArgs& a, const KqDpoly& d, const KqDpolyGeom& g, cudaStream_t st) {
  void* params[] = {(void*)&a, (void*)&d};
  KqPlan pl = {};
  pl.grid = (a.NT + d.TPC - 1) / d.TPC;
  pl.grid_y = a.K;
  pl.block = d.TPC * a.N * a.N;
  pl.smem = g.smem_build;
  return launch(k_dpoly_build<NMAX>, pl, false, st, params);
}
}  // namespace

int kq_launch_dpoly(const KqSweepArgs& a, const KqDpoly& d, const KqDpolyGeom& g,
                    cudaStream_t st) {
  k_dpoly_plan<<<1, 256, 0, st>>>(a, d);
  KQ_CUDA(cudaGetLastError());
  int rc;
  switch (g.nmax) {
    case 4: rc = build_and_zeta<4>(a, d, g, st); break;
    case 8: rc = build_and_zeta<8>(a, d, g, st); break;
    default: rc = build_and_zeta<16>(a, d, g, st); break;
  }
  if (rc) return rc;
  void* params[] = {(void*)&a, (void*)&d};
  KqPlan pl = {};
  pl.grid = 1;
  pl.block = (d.NL + 31) / 32 * 32 + 32;   // consumers + the producer warp
  pl.smem = g.smem_sweep;
  switch (d.C) {
    case 1: return launch(k_dpoly_sweep<1>, pl, false, st, params);
    case 2: return launch(k_dpoly_sweep<2>, pl, false, st, params);
    case 3: return launch(k_dpoly_sweep<3>, pl, false, st, params);
    default: return launch(k_dpoly_sweep<4>, pl, false, st, params);
  }
}

int kq_launch_dpoly_epilogue(const KqSweepArgs& a, const KqDpoly& d, cudaStream_t st) {
  k_dpoly_epilogue<<<1, 256, 0, st>>>(a, d);
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}
