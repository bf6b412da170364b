// Host-side declarations shared by the translation units of
// libkrotov_b200.so.  The kernels are instantiated in several .cu files so
// that they compile in parallel; each file has its own copy of the constant
// Taylor tables (uploaded by its kq_tables_upload_* function).
#pragma once
#include <cuda_runtime.h>

#include <mutex>

#include "../../include/krotov_b200.h"
#include "kq_common.cuh"
#include "kq_small.cuh"   // KqSweepArgs
#include "kq_warp_geom.cuh"

int kq_fail(int code, const char* fmt, ...);

#define KQ_CUDA(call)                                                                 \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess)                                                            \
      return kq_fail(KQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                             \
  } while (0)

struct KqPlan {
  int spec;    // specialised M=2 (L=1) straight-line kernels (kq_spec.cuh)
  int family;  // 0 thread-per-objective, 1 lane-per-row
  int grid, block;
  int grid_y;  // segments of a time-parallel propagation (1 otherwise)
  size_t smem;
  KqWarpGeom geom;
  int rpl;
};


typedef KqPlan Plan;

// Launch helper.  The attribute / occupancy queries are cached per kernel
// instantiation (they cost several microseconds per call and the fused
// iteration kernel is launched once per Krotov iteration).
// kq_set_option("cooperative_launch", 0): kernels that need co-resident CTAs are
// still checked against the occupancy limit but launched with cudaLaunchKernel
extern int g_kq_coop_launch;
// kq_set_option("programmatic_launch", 1) (experimental, needs cooperative_launch 0): launch
// with programmatic stream serialization, so that the next launch's CTAs become resident while
// this grid drains (profiles/micro/launch_gap.cu: kernel boundary 2.9 -> 0.7 us); the kernels
// call griddepcontrol.wait before touching memory, so stream-order semantics are unchanged
extern int g_kq_pdl_launch;

template <typename Kern>
int launch(Kern kern, const Plan& pl, bool cooperative, cudaStream_t st, void** params) {
  // Kern is only the function-pointer TYPE (all sweep kernels share it): the
  // cache is keyed on the pointer value
  struct Entry {
    const void* fn;
    int dev, smem, block, per_sm, sms;
  };
  static Entry cache[128];
  static int n_cache = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  KQ_CUDA(cudaGetDevice(&dev));
  const void* fn = (const void*)kern;
  int c_per_sm = 0, c_sms = 0;
  bool hit = false;
  for (int i = 0; i < n_cache; ++i) {
    const Entry& e = cache[i];
    if (e.fn == fn && e.dev == dev && e.smem == (int)pl.smem && e.block == pl.block) {
      c_per_sm = e.per_sm;
      c_sms = e.sms;
      hit = true;
      break;
    }
  }
  if (!hit) {
    KQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    KQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c_per_sm, kern, pl.block, pl.smem));
    KQ_CUDA(cudaDeviceGetAttribute(&c_sms, cudaDevAttrMultiProcessorCount, dev));
    // a kernel launched with several shared-memory sizes keeps the largest
    // opt-in attribute: drop its older entries so that a smaller size sets it again
    int w = 0;
    for (int i = 0; i < n_cache; ++i)
      if (!(cache[i].fn == fn && cache[i].dev == dev)) cache[w++] = cache[i];
    n_cache = w;
    if (n_cache < 128) cache[n_cache++] = Entry{fn, dev, (int)pl.smem, pl.block, c_per_sm, c_sms};
  }
  if (cooperative) {
    if ((long long)c_per_sm * c_sms < (long long)pl.grid * (pl.grid_y > 1 ? pl.grid_y : 1))
      return kq_fail(KQ_ERR_UNSUPPORTED,
                  "fused sweep needs %d co-resident CTAs but only %d fit (K too large)", pl.grid,
                  c_per_sm * c_sms);
    if (g_kq_coop_launch) {
      KQ_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(pl.grid, pl.grid_y > 1 ? pl.grid_y : 1), dim3(pl.block),
                                          params, pl.smem, st));
    } else if (g_kq_pdl_launch) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(pl.grid, pl.grid_y > 1 ? pl.grid_y : 1);
      cfg.blockDim = dim3(pl.block);
      cfg.dynamicSmemBytes = pl.smem;
      cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      KQ_CUDA(cudaLaunchKernelExC(&cfg, (const void*)kern, params));
    } else {
      KQ_CUDA(cudaLaunchKernel((const void*)kern, dim3(pl.grid, pl.grid_y > 1 ? pl.grid_y : 1),
                               dim3(pl.block), params, pl.smem, st));
    }
  } else {
    KQ_CUDA(cudaLaunchKernel((const void*)kern, dim3(pl.grid, pl.grid_y > 1 ? pl.grid_y : 1),
                             dim3(pl.block), params, pl.smem, st));
  }
  return KQ_OK;
}


#define KQ_DEFINE_TABLES_UPLOAD(name)                                   \
  int name(const KqTables* T) {                                         \
    KQ_CUDA(cudaMemcpyToSymbol(c_kq_tables, T, sizeof(KqTables)));      \
    return KQ_OK;                                                       \
  }

// per-translation-unit entry points
int kq_tables_upload_abi(const KqTables* T);
int kq_tables_upload_small(const KqTables* T);
int kq_tables_upload_spec_prop(const KqTables* T);
int kq_tables_upload_spec_prop_re(const KqTables* T);
int kq_tables_upload_spec_fw2_re(const KqTables* T);
int kq_tables_upload_spec_fw3_re(const KqTables* T);
int kq_tables_upload_spec_fw4_re(const KqTables* T);
int kq_tables_upload_spec_fw2(const KqTables* T);
int kq_tables_upload_spec_fw3(const KqTables* T);
int kq_tables_upload_spec_fw4(const KqTables* T);
int kq_tables_upload_warp0(const KqTables* T);
int kq_tables_upload_warp8(const KqTables* T);
int kq_tables_upload_warp16(const KqTables* T);
int kq_tables_upload_warp32(const KqTables* T);
int kq_tables_upload_picard2(const KqTables* T);
int kq_tables_upload_picard3(const KqTables* T);
int kq_tables_upload_picard4(const KqTables* T);
int kq_tables_upload_dpoly(const KqTables* T);
int kq_tables_upload_csr(const KqTables* T);
int kq_tables_upload_lanes(const KqTables* T);
int kq_tables_upload_sat(const KqTables* T);

int kq_launch_prop_small(const KqSweepArgs& a, const KqPlan& pl, int fsel, cudaStream_t st);
int kq_launch_fwupd_small(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                          cudaStream_t st);
int kq_launch_prop_spec(const KqSweepArgs& a, const KqPlan& pl, int fsel, int nseg,
                        cudaStream_t st);
int kq_launch_prop_spec_re(const KqSweepArgs& a, const KqPlan& pl, int fsel, int nseg,
                           cudaStream_t st);
int kq_launch_fwupd_spec2_re(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                             cudaStream_t st);
int kq_launch_fwupd_spec3_re(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                             cudaStream_t st);
int kq_launch_fwupd_spec4_re(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                             cudaStream_t st);
int kq_launch_fwupd_spec2(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                          cudaStream_t st);
int kq_launch_fwupd_spec3(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                          cudaStream_t st);
int kq_launch_fwupd_spec4(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                          cudaStream_t st);
int kq_launch_picard2(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                            bool real, cudaStream_t st);
int kq_launch_picard3(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                            bool real, cudaStream_t st);
int kq_launch_picard4(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second,
                            bool real, cudaStream_t st);
// row-per-thread CSR kernels for 64 < N <= 1024 (kq_csr.cuh)
struct KqCsr;
int kq_launch_csr(const KqSweepArgs& a, const KqCsr& s, const KqPlan& pl, int fsel, bool update,
                  bool staged, cudaStream_t st);
// one-warp update sweep for few objectives with several controls (kq_lanes.cuh):
// pre-pass + chain
struct KqLanes;
int kq_launch_lanes(const KqSweepArgs& a, const KqLanes& d, int fsel, cudaStream_t st);
int kq_launch_prop_lanes(const KqSweepArgs& a, const KqLanes& d, int fsel, int n_task,
                         cudaStream_t st);
// update sweep for many two-level objectives with a real generator (kq_sat.cuh)
extern int g_kq_sat_min_kpc;   // kq_set_option("sat_min_kpc", v)
int kq_sat_kpc(int K, int sms);
size_t kq_sat_smem(int kpc);
int kq_launch_sat(const KqSweepArgs& a, int sms, cudaStream_t st);
// delta-polynomial update sweep (kq_dpoly.cuh)
struct KqDpoly;
struct KqDpolyGeom {
  int nmax;             // register rows of the build kernel (4, 8, 16, 32)
  size_t smem_build, smem_sweep;
};
int kq_launch_dpoly(const KqSweepArgs& a, const KqDpoly& d, const KqDpolyGeom& g, cudaStream_t st);
int kq_launch_dpoly_epilogue(const KqSweepArgs& a, const KqDpoly& d, int fallback_in_stream,
                             cudaStream_t st);
int kq_launch_warp0(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool update,
                    cudaStream_t st);
int kq_launch_warp8(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool update,
                    cudaStream_t st);
int kq_launch_warp16(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool update,
                     cudaStream_t st);
int kq_launch_warp32(const KqSweepArgs& a, const KqPlan& pl, int fsel, bool second, bool update,
                     cudaStream_t st);
