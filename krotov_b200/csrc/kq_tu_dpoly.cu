// Delta-polynomial Krotov iteration (kq_dpoly.cuh): launches.
#include <algorithm>

#include "kq_host.cuh"
#include "kq_dpoly.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_dpoly)

namespace {
int round32(int v) { return (v + 31) / 32 * 32; }

template <int NMAX>
int build_and_backward(const KqSweepArgs& a, const KqDpoly& d, const KqDpolyGeom& g,
                       cudaStream_t st) {
  void* params[] = {(void*)&a, (void*)&d};
  KqPlan pl = {};
  pl.grid = (a.NT + 2 * KQ_DP_PAD + d.TPC - 1) / d.TPC;
  pl.grid_y = a.K;
  pl.block = d.TPC * a.N * a.N;
  pl.smem = g.smem_build;
  int rc = launch(k_dp_build<NMAX>, pl, false, st, params);
  if (rc) return rc;
  const int NN = a.N * a.N;
  if (d.chain) {
    KqPlan ps = {};
    ps.grid = d.nseg;
    ps.grid_y = a.K;
    ps.block = std::max(128, round32(NN));   // extra threads help staging the records
    ps.smem = ((size_t)3 * NN + (size_t)2 * d.estage_cap) * sizeof(cplx) + (size_t)d.seg_len * 8;
    rc = launch(k_dp_segprod, ps, false, st, params);
    if (rc) return rc;
  }
  KqPlan pe = {};
  pe.grid = d.nseg;
  pe.grid_y = a.K;
  pe.block = std::max(128, round32(a.N * d.R2));
  pe.smem = ((size_t)3 * a.N + (size_t)2 * d.chain_bs * NN + (size_t)2 * d.estage_cap + NN) * sizeof(cplx) +
            (size_t)d.seg_len * 8;
  rc = launch(k_dp_expand<NMAX>, pe, false, st, params);
  if (rc) return rc;
  return KQ_OK;
}

template <bool QONE, bool ONEWARP, int MAXT, bool PF>
int sweep_c(const KqSweepArgs& a, const KqDpoly& d, const KqDpolyGeom& g, cudaStream_t st) {
  void* params[] = {(void*)&a, (void*)&d};
  KqPlan pl = {};
  pl.grid = 1;
  pl.block = round32(d.NL) + 32;   // consumers + the producer warp
  pl.smem = g.smem_sweep;
  switch (d.C) {
    case 1: return launch(k_dp_sweep<1, QONE, ONEWARP, MAXT, PF>, pl, false, st, params);
    case 2: return launch(k_dp_sweep<2, QONE, ONEWARP, MAXT, PF>, pl, false, st, params);
    default: return launch(k_dp_sweep<4, QONE, ONEWARP, MAXT, PF>, pl, false, st, params);
  }
}
template <bool ONEWARP, int MAXT, bool PF>
int sweep(const KqSweepArgs& a, const KqDpoly& d, const KqDpolyGeom& g, cudaStream_t st) {
  return d.Q == 1 ? sweep_c<true, ONEWARP, MAXT, PF>(a, d, g, st)
                  : sweep_c<false, ONEWARP, MAXT, PF>(a, d, g, st);
}
}  // namespace

int kq_launch_dpoly(const KqSweepArgs& a, const KqDpoly& d, const KqDpolyGeom& g,
                    cudaStream_t st) {
  k_dp_plan<<<1, 256, 0, st>>>(a, d);
  KQ_CUDA(cudaGetLastError());
  int rc;
  switch (g.nmax) {
    case 4: rc = build_and_backward<4>(a, d, g, st); break;
    case 8: rc = build_and_backward<8>(a, d, g, st); break;
    default: rc = build_and_backward<16>(a, d, g, st); break;
  }
  if (rc) return rc;
  if (d.NL <= 32) return sweep<true, 64, true>(a, d, g, st);
  if (d.NL <= 224) return sweep<false, 256, true>(a, d, g, st);
  return sweep<false, KQ_DP_MAXLANES + 32, false>(a, d, g, st);
}

int kq_launch_dpoly_epilogue(const KqSweepArgs& a, const KqDpoly& d, int fallback_in_stream,
                             cudaStream_t st) {
  k_dp_epilogue<<<1, 256, 0, st>>>(a, d, fallback_in_stream);
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}
