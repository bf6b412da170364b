// C ABI of libkrotov_b200.so: argument checking, launch planning and kernel
// dispatch for the Krotov sweep hot path (see include/krotov_b200.h for the
// reference call sites each entry point replaces).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "kq_host.cuh"
#include "kq_dpoly_geom.cuh"
#include "kq_csr.cuh"
#include "kq_lanes_geom.cuh"

int g_kq_coop_launch = 1;
int g_kq_pdl_launch = 0;

namespace {

thread_local std::string g_err;

}  // namespace

int kq_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_abi)

namespace {

#define fail kq_fail

constexpr int kMaxDevices = 64;
constexpr int kMaxBlocks = KQ_MAX_BLOCKS;
constexpr int kMaxWorld = KQ_MAX_WORLD;
constexpr size_t kStatusBytes = 256;
constexpr size_t kSmemBudget = 200 * 1024;

struct DeviceInfo {
  bool ready = false;
  int sms = 0;
  int max_smem_optin = 0;
  int coop = 0;
};
DeviceInfo g_dev[kMaxDevices];
std::mutex g_mu;

// Per-device scratch for the time-parallel propagation (segment propagators
// and boundary states).  Grown on demand; calls on one device are expected to
// be stream-ordered with respect to each other.
struct Scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};
Scratch g_scratch[kMaxDevices][3];   // slot 0: time-parallel propagation, 1: dpoly records,
                                     // 2: records of the one-warp update sweep (kq_lanes.cuh)
bool g_disable_segments = false;   // kq_set_option("time_parallel", 0)
int g_picard = 1;                  // kq_set_option("picard", 0|1|2): off / auto / forced
int g_picard_timing = 0;
int g_picard_history = 1;          // kq_set_option("picard_history", 0|1): update-history hint
int g_picard_maxit = 64;           // kq_set_option("picard_maxit", n)
int g_picard_rtol_e15 = 20;        // kq_set_option("picard_rtol_e15", v): fixed point accepted at v * 1e-15
int g_dpoly = 1;                   // kq_set_option("dpoly", 0|1|2): off / auto / wherever it fits
int g_dpoly_debug = 0;
int g_lanes = 1;                    // kq_set_option("lanes", 0): without the one-warp update sweep (kq_lanes.cuh)
int g_sat = 1;                      // kq_set_option("sat", 0): without the many-objective update sweep (kq_sat.cuh)
int g_small_rows = 0;               // kq_set_option("small_rows", 0): thread-per-objective kernels for few generic objectives too             // kq_set_option("dpoly_debug", 1): no sequential kernel behind it
constexpr int kPicMaxBlocks = 148;   // CTAs of the time-parallel fused sweep (one per SM)
constexpr int kPicMaxItCap = 1000;
constexpr int kPicMaxWindows = 64;   // time windows of one windowed update sweep

int get_scratch(int dev, size_t bytes, void** out, int slot = 0) {
  std::lock_guard<std::mutex> lock(g_mu);
  Scratch& sc = g_scratch[dev][slot];
  if (sc.bytes < bytes) {
    if (sc.ptr) KQ_CUDA(cudaFree(sc.ptr));   // synchronises the device
    sc.ptr = nullptr;
    sc.bytes = 0;
    KQ_CUDA(cudaMalloc(&sc.ptr, bytes));
    sc.bytes = bytes;
  }
  *out = sc.ptr;
  return KQ_OK;
}

// Taylor degree per binade: smallest m with  b^(m+1)/(m+1)! <= 2^-56,
// b = min(1, 2^-bin) the largest scaled norm in the bin.
void build_tables(KqTables& T) {
  const double tol = std::ldexp(1.0, -56);
  for (int bin = 0; bin < KQ_TAYLOR_BINS; ++bin) {
    const double b = std::ldexp(1.0, -bin);
    double term = b;  // b^1/1!
    int m = 0;        // term = b^(m+1)/(m+1)!
    while (term > tol && m < KQ_TAYLOR_MAXM) {
      ++m;
      term *= b / (double)(m + 1);
    }
    T.m_of_bin[bin] = m < 1 ? 1 : m;
  }
  for (int i = 0; i < 4 * KQ_TAYLOR_BINS; ++i) {
    // xs = 2^-(bin+1) (1 + f), the two leading mantissa bits of f select the quarter
    const double b = std::min(1.0, std::ldexp(1.0 + (double)(i % 4 + 1) / 4.0, -(i / 4 + 1)));
    double term = b;
    int m = 0;
    while (term > tol && m < KQ_TAYLOR_MAXM) {
      ++m;
      term *= b / (double)(m + 1);
    }
    T.m_fine[i] = (unsigned char)(m < 1 ? 1 : m);
  }
  T.inv[0] = 0.0;
  for (int j = 1; j <= KQ_TAYLOR_MAXM; ++j) T.inv[j] = 1.0 / (double)j;
  double f = 1.0;
  for (int n = 0; n < KQ_INVFACT_N; ++n) {
    if (n > 1) f *= (double)n;
    T.invfact[n] = 1.0 / f;
  }
}

int device_init(int* dev_out) {
  int dev = 0;
  KQ_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(KQ_ERR_ARG, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_mu);
  DeviceInfo& d = g_dev[dev];
  if (!d.ready) {
    KqTables T;
    build_tables(T);
    // every translation unit holds its own constant copy of the tables
    int (*const uploads[])(const KqTables*) = {
        kq_tables_upload_abi,      kq_tables_upload_small,    kq_tables_upload_spec_prop,
        kq_tables_upload_spec_fw2, kq_tables_upload_spec_fw3, kq_tables_upload_spec_fw4,
        kq_tables_upload_spec_prop_re, kq_tables_upload_spec_fw2_re,
        kq_tables_upload_spec_fw3_re, kq_tables_upload_spec_fw4_re,
        kq_tables_upload_warp0,    kq_tables_upload_warp8,    kq_tables_upload_warp16,
        kq_tables_upload_warp32,   kq_tables_upload_picard2,  kq_tables_upload_picard3,
        kq_tables_upload_picard4,  kq_tables_upload_dpoly,    kq_tables_upload_csr,
        kq_tables_upload_lanes,    kq_tables_upload_sat};
    for (auto up : uploads) {
      const int rc = up(&T);
      if (rc) return rc;
    }
    KQ_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
    KQ_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KQ_CUDA(cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, dev));
    d.ready = true;
  }
  *dev_out = dev;
  return KQ_OK;
}

int check_problem(const kq_problem* p) {
  if (!p) return fail(KQ_ERR_ARG, "problem is NULL");
  if (p->K < 1 || p->N < 1 || p->NT < 1 || p->L < 0 || p->M < 1)
    return fail(KQ_ERR_ARG, "invalid sizes K=%d N=%d NT=%d L=%d M=%d", p->K, p->N, p->NT, p->L, p->M);
  if (p->L > KQ_LMAX) return fail(KQ_ERR_UNSUPPORTED, "L=%d pulses > %d", p->L, KQ_LMAX);
  if (p->N > 1024)
    return fail(KQ_ERR_UNSUPPORTED, "state length N=%d > 1024 is not built", p->N);
  if (p->N > 64 || (p->sparse && !p->ops)) {
    if (!p->sparse || !p->sparse->row_ptr || !p->sparse->mat_off || !p->sparse->col || !p->sparse->val)
      return fail(KQ_ERR_ARG, "N=%d > 64 needs the sparse form of the matrices (kq_problem.sparse)",
                  p->N);
    if (!p->term2pulse || !p->op_norm || !p->dt)
      return fail(KQ_ERR_ARG, "problem has NULL time / mapping arrays");
    return KQ_OK;
  }
  if (!p->ops || !p->ops_adj || !p->term2pulse || !p->op_norm || !p->dt)
    return fail(KQ_ERR_ARG, "problem has NULL operator/time arrays");
  return KQ_OK;
}

int round_up(int v, int q) { return (v + q - 1) / q * q; }

// update = fused sweep (objectives should share a CTA); otherwise spread
// independent objectives over the SMs.
int make_plan(const kq_problem* p, bool update, bool second, int sms, Plan& pl,
              bool force_rows = false) {
  const int K = p->K, N = p->N, M = p->M, L = p->L, NN = N * N;
  std::memset(&pl, 0, sizeof pl);
  if (N > 64 || (p->sparse && !p->ops)) {
    // row-per-thread CSR family (kq_csr.cuh): one CTA per objective
    pl.family = 2;
    pl.block = round_up(N, 32);
    pl.grid = K;
    pl.smem = (size_t)(2 * KQ_LMAX * 32 + 2 * KQ_LMAX + ((M + 1) & ~1)) * sizeof(double) +
              (size_t)2 * N * sizeof(cplx);
    (void)second;
    (void)sms;
    return KQ_OK;
  }
  // Few objectives with several terms / controls (the Lambda systems of notebooks 02/03/08:
  // K = 1..5, N = 3, four controls): the thread-per-objective kernels would run one thread
  // per CTA; the lane-per-row family spreads rows and keeps its row of A in registers.
  const bool few_generic =
      force_rows || (N <= 4 && !(M == 2 && (!update || L == 1)) && K <= 8 && g_small_rows);
  if (N <= 4 && M <= KQ_MMAX_SMALL && !few_generic) {
    pl.family = 0;
    pl.spec = (N >= 2 && M == 2 && (!update || L == 1)) ? 1 : 0;
    size_t per_thread = (size_t)(M + (update ? L : 0)) * NN * sizeof(cplx);
    size_t fixed = (2 * KQ_LMAX * 32 + 2 * KQ_LMAX) * sizeof(double);
    if (pl.spec) {
      per_thread = (N == 4) ? (size_t)2 * NN * sizeof(cplx) : 0;
      if (update) {
        per_thread += (size_t)KQ_RING * N * sizeof(cplx) * (second ? 2 : 1);
        fixed = 72 * sizeof(double) + KQ_RING * sizeof(uint64_t) +
                5 * KQ_NTC * sizeof(double) + KQ_NTC;
      } else {
        fixed = 2 * KQ_NTC * sizeof(double) + KQ_NTC + 32 * sizeof(double);
      }
    }
    int cap = per_thread ? (int)((kSmemBudget - fixed) / per_thread) / 32 * 32 : 1024;
    const int maxbt = update ? KQ_SMALL_MAXBT(N) : 256;
    if (cap > maxbt) cap = maxbt;
    if (cap < 32) return fail(KQ_ERR_UNSUPPORTED, "generator terms do not fit shared memory");
    int bt;
    if (update) {
      bt = std::min(cap, round_up(K, 32));
    } else {
      bt = std::min(cap, std::max(32, round_up((K + sms - 1) / sms, 32)));
    }
    pl.block = bt;
    pl.grid = (K + bt - 1) / bt;
    pl.smem = per_thread * bt + fixed;
    return KQ_OK;
  }
  pl.family = 1;
  int R = 2;
  while (R < N && R < 32) R <<= 1;
  pl.rpl = (N + 31) / 32;
  if (pl.rpl < 1) pl.rpl = 1;
  KqWarpGeom& g = pl.geom;
  g.R = R;
  g.G = 32 / R;
  const size_t fixed = (2 * KQ_LMAX * 32 + 2 * KQ_LMAX) * sizeof(double);
  // state double buffer: padded to the register-row capacity (8/16/32) for
  // N <= 32 (kq_warp.cuh MODE >= 8), else to a multiple of 4
  const int npad = (N <= 8) ? 8 : (N <= 16 ? 16 : (N <= 32 ? 32 : ((N + 3) & ~3)));
  const size_t base = (size_t)NN + 2 * npad + (M + 1) / 2;
  const size_t with_mu = base + (update ? (size_t)L * NN : 0);
  const size_t with_all = with_mu + (size_t)M * NN;
  size_t stride;
  if (fixed + with_all * sizeof(cplx) * g.G <= kSmemBudget) {
    g.terms_in_smem = 1;
    g.mu_in_smem = update ? 1 : 0;
    stride = with_all;
  } else if (fixed + with_mu * sizeof(cplx) * g.G <= kSmemBudget) {
    g.terms_in_smem = 0;
    g.mu_in_smem = update ? 1 : 0;
    stride = with_mu;
  } else if (fixed + base * sizeof(cplx) * g.G <= kSmemBudget) {
    g.terms_in_smem = 0;
    g.mu_in_smem = 0;
    stride = base;
  } else {
    return fail(KQ_ERR_UNSUPPORTED, "N=%d does not fit shared memory", N);
  }
  g.obj_stride = (int)stride;
  const size_t per_warp = stride * sizeof(cplx) * g.G;
  int max_warps = (int)((kSmemBudget - fixed) / per_warp);
  if (max_warps > 16) max_warps = 16;  // kernel is built for <= 512 threads
  if (N > 8 && N <= 16 && max_warps > 8) max_warps = 8;   // register-row modes are built
  if (N > 16 && N <= 32 && max_warps > 4) max_warps = 4;  // for 256 / 128 threads
  if (max_warps < 1) max_warps = 1;
  const int warps_needed = (K + g.G - 1) / g.G;
  int wpb;
  if (update) {
    wpb = std::min(max_warps, warps_needed);
  } else {
    wpb = std::min(max_warps, std::max(1, (warps_needed + sms - 1) / sms));
  }
  pl.block = wpb * 32;
  pl.grid = (warps_needed + wpb - 1) / wpb;
  pl.smem = fixed + per_warp * wpb;
  return KQ_OK;
}

// Geometry of the time-parallel fused sweep (kq_picard.cuh): Q objectives per
// CTA, TC chunks of W steps per objective.  Returns false if the problem is
// outside what that family handles (the sequential kernels take over).
struct PicPlan {
  int Q, TC, W, lw, Wc, lwc, grid, block, stride;
  size_t smem;
};
// slots per mailbox / per owner region: grid * 2^lwc + grid <= 2 NT + 19 * 148
// slots of a rank's exchange buffer that precede the cross-GPU region of the time-parallel
// iteration: per-step slots of the sequential sweep | barrier flags
size_t xg_offset_slots() { return (size_t)2 * kMaxWorld * KQ_LMAX + kMaxWorld; }
size_t pic_stride(const kq_problem* p) { return (size_t)2 * round_up(p->NT, 64) + 2880; }
int pic_hist_ld(const kq_problem* p) { return round_up(p->NT, 16); }
// `K_plan` (default p->K): number of objectives the geometry is made for -- with objectives
// sharded over GPUs every rank plans for the largest block, so that all ranks run the same
// grid and the same time slices.
bool picard_plan(const kq_problem* p, int sms, PicPlan& pp, int K_plan = 0) {
  const int K = K_plan > 0 ? K_plan : p->K, N = p->N, NT = p->NT, NN = N * N;
  if (N < 2 || N > 4 || p->M != 2 || p->L != 1) return false;
  const int gmax = std::min(sms, kPicMaxBlocks);
  const int Q = (K + gmax - 1) / gmax;
  if (Q > 8) return false;
  const int tcmax = (256 / Q) / 32 * 32;
  // chunk length: the smallest power of two that covers the grid with <= tcmax
  // chunks; then as few chunks as needed
  int lw = 0;
  while (((NT + (1 << lw) - 1) >> lw) > tcmax) ++lw;
  const int W = 1 << lw;
  const int TC = round_up((NT + W - 1) / W, 32);
  const size_t NTP = (size_t)TC * W;
  const int grid = (K + Q - 1) / Q;
  const int Wc = round_up((NT + grid - 1) / grid, 2);
  int lwc = 3;
  while ((1 << lwc) < Wc) ++lwc;
  if (((size_t)grid << lwc) + grid > pic_stride(p)) return false;
  const size_t smem = 256 * sizeof(double) + 2 * NTP * sizeof(double) + (size_t)Q * NTP * sizeof(double) +
                      (size_t)4 * Wc * sizeof(double) + (size_t)Q * NTP * N * sizeof(cplx) +
                      (size_t)2 * Q * 8 * NN * sizeof(cplx) + (size_t)Q * 4 * NN * sizeof(cplx);
  if (smem > kSmemBudget) return false;
  pp.Q = Q;
  pp.TC = TC;
  pp.W = W;
  pp.lw = lw;
  pp.Wc = Wc;
  pp.lwc = lwc;
  pp.grid = grid;
  pp.block = Q * TC;
  pp.stride = (int)pic_stride(p);
  pp.smem = smem;
  return true;
}

size_t dpoly_header_offset(const kq_problem* p) {
  size_t bytes = kStatusBytes + (size_t)2 * kMaxBlocks * KQ_LMAX * sizeof(KqSlot);
  if (p && p->NT > 0) {   // slots of the time-parallel fused sweep: part | eps | ga
    bytes += (size_t)2 * kPicMaxBlocks * pic_stride(p) * sizeof(KqSlot);
    bytes += 64 + (size_t)4 * pic_hist_ld(p) * sizeof(double);   // update history: header | ring
  }
  return (bytes + 63) / 64 * 64;
}

// ---- delta-polynomial Krotov iteration (kq_dpoly.cuh): few objectives, one control ----
struct DpPlan {
  KqDpoly d;
  KqDpolyGeom g;
};
bool dpoly_plan(const kq_problem* p, DpPlan& dp) {
  const int K = p->K, N = p->N;
  if (p->M != 2 || p->L != 1 || N < 2 || N > 16 || p->NT < 2) return false;
  int Npad = 2;
  while (Npad < N) Npad <<= 1;
  const int NR = N + 1;
  // lanes per row Q (C = Npad / Q columns per lane): the largest Q that keeps all lanes in
  // ONE warp (the chain then runs on shuffles alone), else the largest that fits the CTA
  int Q = 0;
  for (int q = std::min(Npad, 8); q >= 1; q >>= 1)
    if (Npad / q <= KQ_DP_CMAX && (long long)K * NR * q <= 32) {
      Q = q;
      break;
    }
  if (!Q)
    for (int q = std::min(Npad, 8); q >= 1; q >>= 1)
      if (Npad / q <= KQ_DP_CMAX && (long long)K * NR * q <= KQ_DP_MAXLANES) {
        Q = q;
        break;
      }
  if (!Q) return false;
  int C = Npad / Q;
  if (C == 3) return false;   // cannot happen (powers of two)
  const long long NL = (long long)K * NR * Q;   // rows of E plus the zeta row
  const size_t rec_stride = (size_t)C * NL * ((KQ_DP_JMAX + 1) | 1) + 2;
  // sweep kernel: barriers | exchange buffers [2][XS] | first overlap | ring of record chunks
  const size_t nlp = (size_t)round_up((int)NL, 32);
  const size_t XS = (size_t)2 * K * Npad + 2 * K + 2 * nlp;
  const size_t fixed = 2 * KQ_DP_RINGMAX * 8 + 2 * XS * 8 + (size_t)((K * N + 1) & ~1) * 8;
  const size_t budget = 220 * 1024;
  // two stages of two records of the highest degree must fit
  if (fixed + 4 * rec_stride * 16 > budget) return false;
  const size_t ring = (budget - fixed) / 16;
  // build kernel: one thread per matrix element, polynomial double-buffered
  const size_t per = (size_t)2 * (KQ_DP_JMAX + 1) * N * N * 16;
  const int TPC = std::max(1, std::min(256 / (N * N), (int)(100 * 1024 / per)));
  if ((size_t)TPC * per > kSmemBudget) return false;
  std::memset(&dp, 0, sizeof dp);
  dp.d.rec_stride = (int)rec_stride;
  dp.d.NL = (int)NL;
  dp.d.Q = Q;
  dp.d.C = C;
  dp.d.Npad = Npad;
  dp.d.R2 = Npad;
  dp.d.chain_bs = std::max(1, std::min(32, 1024 / (N * N)));
  // one record of the highest degree must fit a staging buffer; small ones: about 24 KB
  dp.d.estage_cap = std::max((KQ_DP_JMAX + 1) * N * N, 1536);
  dp.d.TPC = TPC;
  dp.d.ring = (int)ring;
  // backward sweep: segments of about 0.46 sqrt(NT) steps (the expand kernel chains over
  // the later segments, then walks through its own)
  int seg_len = std::max(4, (int)std::ceil(0.46 * std::sqrt((double)p->NT)));
  while ((p->NT + seg_len - 1) / seg_len > KQ_DP_SEGMAX) ++seg_len;
  dp.d.seg_len = seg_len;
  dp.d.nseg = (p->NT + seg_len - 1) / seg_len;
  dp.g.nmax = N <= 4 ? 4 : (N <= 8 ? 8 : 16);
  dp.g.smem_build = (size_t)TPC * per;
  dp.g.smem_sweep = fixed + ring * 16;
  return true;
}
// The delta-polynomial iteration is the only fast path for N > 4; for N <= 4 the caller
// asks for it (kq_problem.update_sweep = 1) when the problem is strongly coupled or has few
// objectives -- the time-parallel fixed point needs many rounds, or leaves most of the GPU
// idle, there.
bool dpoly_preferred(const kq_problem* p) {
  if (g_dpoly == 2) return true;
  return g_dpoly == 1 && (p->N > 4 || p->update_sweep == 1);
}
// workspace behind the header: anchor pulse | step records (kept between calls)
size_t dpoly_anchor_bytes(const kq_problem* p) { return ((size_t)p->NT * sizeof(double) + 63) / 64 * 64; }
size_t dpoly_rec_bytes(const kq_problem* p, const DpPlan& dp) {
  return ((size_t)(p->NT + 2 * KQ_DP_PAD) * dp.d.rec_stride * sizeof(cplx) + 255) / 256 * 256;
}
size_t dpoly_workspace_bytes(const kq_problem* p) {
  DpPlan dp;
  if (!p || p->NT < 1 || !dpoly_plan(p, dp)) return 0;
  return dpoly_anchor_bytes(p) + dpoly_rec_bytes(p, dp);
}
// per-call scratch (device-global, stream-ordered): segment propagators | X | chi | phi(T) |
// chi norms | tau sum
struct DpScratch {
  cplx *segP, *X, *chi, *phiT, *tsum;
  double* norms;
};
int dpoly_scratch(const kq_problem* p, const DpPlan& dp, int dev, DpScratch& s) {
  const int K = p->K, N = p->N, NT = p->NT;
  const size_t pb = ((size_t)dp.d.nseg * K * N * N * sizeof(cplx) + 255) / 256 * 256;
  const size_t xb = ((size_t)(NT + 1) * K * N * sizeof(cplx) + 255) / 256 * 256;
  const size_t sb = ((size_t)K * N * sizeof(cplx) + 255) / 256 * 256;
  const size_t nb = ((size_t)K * sizeof(double) + 255) / 256 * 256;
  void* base = nullptr;
  int rc = get_scratch(dev, pb + xb + 2 * sb + nb + 256, &base, 1);
  if (rc) return rc;
  char* cur = static_cast<char*>(base);
  s.segP = reinterpret_cast<cplx*>(cur);
  cur += pb;
  s.X = reinterpret_cast<cplx*>(cur);
  cur += xb;
  s.chi = reinterpret_cast<cplx*>(cur);
  cur += sb;
  s.phiT = reinterpret_cast<cplx*>(cur);
  cur += sb;
  s.norms = reinterpret_cast<double*>(cur);
  cur += nb;
  s.tsum = reinterpret_cast<cplx*>(cur);
  return KQ_OK;
}
// plan | build | (segment products) | expand | sweep on `st`.  The caller has set dp.d.chain,
// X, chi, norms, segP; `a` is the argument block of the update sweep (a.epoch set).
int launch_dpoly(const kq_problem* p, const KqSweepArgs& a, DpPlan& dp, void* workspace,
                 cudaStream_t st) {
  char* w = static_cast<char*>(workspace) + dpoly_header_offset(p);
  dp.d.hdr = reinterpret_cast<KqDpHeader*>(w);
  static_assert(sizeof(KqDpHeader) <= 128, "header area");
  dp.d.anchor = reinterpret_cast<double*>(w + 128);
  dp.d.rec = reinterpret_cast<cplx*>(w + 128 + dpoly_anchor_bytes(p));
  return kq_launch_dpoly(a, dp.d, dp.g, st);
}

// Fill the time-parallel family's launch arguments and launch it.
int launch_picard(const kq_problem* p, KqSweepArgs b, const PicPlan& pp, void* workspace,
                  uint32_t epoch, bool second, cudaStream_t st, int window = 0) {
  b.epoch = epoch ? epoch : 1u;
  b.pic_window = window;
  b.pic_Q = pp.Q;
  b.pic_TC = pp.TC;
  b.pic_W = pp.W;
  b.pic_lw = pp.lw;
  b.pic_Wc = pp.Wc;
  b.pic_lwc = pp.lwc;
  b.pic_stride = pp.stride;
  b.pic_maxit = std::min(g_picard_maxit, p->NT + 1);
  b.pic_rtol = (double)g_picard_rtol_e15 * 1e-15;
  b.pic_timing = g_picard_timing;
  // one tag range per (call, time window)
  b.tag_base = (epoch * (uint32_t)kPicMaxWindows + (uint32_t)window) * (uint32_t)(kPicMaxItCap + 2);
  b.status = reinterpret_cast<int*>(workspace);
  char* base = static_cast<char*>(workspace) + kStatusBytes +
               (size_t)2 * kMaxBlocks * KQ_LMAX * sizeof(KqSlot);
  b.pic_part = reinterpret_cast<KqSlot*>(base);
  b.pic_eps = b.pic_part + (size_t)kPicMaxBlocks * pp.stride;
  if (b.world > 1) {
    b.pic_xg_off = xg_offset_slots();
    b.pic_xg_buf = (size_t)b.world * pp.stride;
  }
  if (b.pic_bw && window == 0 && g_picard_history) {
    // whole-iteration calls keep the last updates in the workspace (first-iterate hint)
    char* hist = reinterpret_cast<char*>(b.pic_eps + (size_t)kPicMaxBlocks * pp.stride);
    b.pic_hist_hdr = reinterpret_cast<unsigned long long*>(hist);
    b.pic_hist = reinterpret_cast<double*>(hist + 64);
    b.pic_hist_ld = pic_hist_ld(p);
  }
  Plan ppl;
  std::memset(&ppl, 0, sizeof ppl);
  ppl.grid = pp.grid;
  ppl.block = pp.block;
  ppl.smem = pp.smem;
  const int fsel = p->is_super ? 2 : 0;
  const bool real = p->real_ops && !p->is_super;
  switch (p->N) {
    case 2: return kq_launch_picard2(b, ppl, fsel, second, real, st);
    case 3: return kq_launch_picard3(b, ppl, fsel, second, real, st);
    default: return kq_launch_picard4(b, ppl, fsel, second, real, st);
  }
}

KqSweepArgs base_args(const kq_problem* p) {
  KqSweepArgs a;
  std::memset(&a, 0, sizeof a);
  a.K = p->K;
  a.N = p->N;
  a.NT = p->NT;
  a.L = p->L;
  a.M = p->M;
  a.is_super = p->is_super;
  a.mu = reinterpret_cast<const cplx*>(p->mu);
  a.term2pulse = p->term2pulse;
  a.op_norm = p->op_norm;
  a.dt = p->dt;
  a.shape = p->shape;
  a.lambda_a = p->lambda_a;
  a.world = 1;
  a.k_lo = 0;
  a.k_cnt = p->K;
  a.seg_len = 0;
  a.seg_pass = 0;
  return a;
}

int launch_warp(const KqSweepArgs& a, const Plan& pl, int fsel, bool second, bool update,
                cudaStream_t st) {
  const bool allsm = pl.geom.terms_in_smem && (!update || pl.geom.mu_in_smem);
  if (pl.rpl == 1 && allsm) {
    if (a.N <= 8) return kq_launch_warp8(a, pl, fsel, second, update, st);
    if (a.N <= 16) return kq_launch_warp16(a, pl, fsel, second, update, st);
    if (a.N <= 32) return kq_launch_warp32(a, pl, fsel, second, update, st);
  }
  return kq_launch_warp0(a, pl, fsel, second, update, st);
}

// Boundary states of the segments for the lane-per-row family (any N <= 64):
// B[0] = state0, B[q+1] = P_q B[q].  One CTA per objective, one thread per row;
// the next segment propagator's row is loaded into registers (NC columns) while
// the current product is formed.
template <int NC>
__global__ void k_seg_chain_rows(const KqSweepArgs a, int nseg) {
  extern __shared__ __align__(16) unsigned char chain_smem[];
  cplx* sb = reinterpret_cast<cplx*>(chain_smem);   // [2][N]
  const int N = a.N, K = a.K, NN = N * N;
  const int k = a.k_lo + blockIdx.x, r = threadIdx.x;
  const bool act = r < N;
  const cplx* __restrict__ segP = a.seg_P;
  cplx* __restrict__ segB = a.seg_B;
  if (act) {
    sb[r] = a.state0[(size_t)k * N + r];
    segB[((size_t)0 * K + k) * N + r] = sb[r];
  }
  cplx Pn[NC > 0 ? NC : 1];
  if (NC > 0 && act) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
      Pn[c] = (c < N) ? segP[((size_t)0 * K + k) * NN + (size_t)c * N + r] : c_zero();
  }
  __syncthreads();
  int p = 0;
  for (int q = 0; q < nseg; ++q) {
    cplx acc0 = c_zero(), acc1 = c_zero(), acc2 = c_zero(), acc3 = c_zero();
    if (act) {
      if (NC > 0) {
        cplx Pc[NC > 0 ? NC : 1];
#pragma unroll
        for (int c = 0; c < NC; ++c) Pc[c] = Pn[c];
        if (q + 1 < nseg) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            Pn[c] = (c < N) ? segP[((size_t)(q + 1) * K + k) * NN + (size_t)c * N + r] : c_zero();
        }
#pragma unroll
        for (int c = 0; c < NC; c += 4) {
          if (c < N) acc0 = c_fma(Pc[c], sb[p * N + c], acc0);
          if (c + 1 < N) acc1 = c_fma(Pc[c + 1], sb[p * N + c + 1], acc1);
          if (c + 2 < N) acc2 = c_fma(Pc[c + 2], sb[p * N + c + 2], acc2);
          if (c + 3 < N) acc3 = c_fma(Pc[c + 3], sb[p * N + c + 3], acc3);
        }
      } else {
        const cplx* P = segP + ((size_t)q * K + k) * NN + r;
        int c = 0;
        for (; c + 3 < N; c += 4) {
          acc0 = c_fma(P[(size_t)c * N], sb[p * N + c], acc0);
          acc1 = c_fma(P[(size_t)(c + 1) * N], sb[p * N + c + 1], acc1);
          acc2 = c_fma(P[(size_t)(c + 2) * N], sb[p * N + c + 2], acc2);
          acc3 = c_fma(P[(size_t)(c + 3) * N], sb[p * N + c + 3], acc3);
        }
        for (; c < N; ++c) acc0 = c_fma(P[(size_t)c * N], sb[p * N + c], acc0);
      }
      const cplx acc = c_add(c_add(acc0, acc1), c_add(acc2, acc3));
      sb[(p ^ 1) * N + r] = acc;
      segB[((size_t)(q + 1) * K + k) * N + r] = acc;
    }
    __syncthreads();
    p ^= 1;
  }
}

KqCsr csr_of(const kq_problem* p) {
  KqCsr s;
  s.row_ptr = p->sparse->row_ptr;
  s.mat_off = reinterpret_cast<const long long*>(p->sparse->mat_off);
  s.col = p->sparse->col;
  s.val = reinterpret_cast<const cplx*>(p->sparse->val);
  s.col16 = p->sparse->col16;
  s.code16 = p->sparse->code16;
  s.dict = reinterpret_cast<const cplx*>(p->sparse->dict);
  s.n_dict = p->sparse->n_dict;
  return s;
}
// Shared memory of the staged variant (dictionary + row pointers + packed non-zeros behind
// the plan's base size); 0 if the matrices of one objective do not fit.
size_t csr_stage_bytes(const kq_problem* p, bool update, size_t base) {
  const kq_sparse* sp = p->sparse;
  const int nnz = update ? sp->stage_nnz_update : sp->stage_nnz_prop;
  if (!sp->col16 || !sp->code16 || !sp->dict || sp->n_dict < 1 || nnz < 1) return 0;
  const int nloc = update ? p->M + p->L : p->M;
  const size_t extra = (size_t)sp->n_dict * sizeof(cplx) + (size_t)nloc * (p->N + 1) * 4 + (size_t)nnz * 4;
  return base + extra <= (size_t)220 * 1024 ? extra : 0;
}

// Geometry of the entries-in-registers propagation kernels (kq_lanes.cuh): N <= 32, at most
// KQ_LN_NZMAX non-zeros per row (kq_problem.row_nnz; 0 = unknown: dense rows).
bool prop_lanes_plan(const kq_problem* p, KqLanes& ln) {
  if (!g_lanes || p->N < 2 || p->N > 32 || p->L > KQ_LN_LMAX || p->M > KQ_MMAX_SMALL || !p->ops ||
      !p->ops_adj)
    return false;
  const int nnz = p->row_nnz > 0 ? p->row_nnz : p->N;
  if (nnz > KQ_LN_NZMAX) return false;
  int NP = 2;
  while (NP < p->N) NP <<= 1;
  std::memset(&ln, 0, sizeof ln);
  ln.NP = NP;
  ln.G = 32 / NP;
  ln.W = KQ_LN_WMAX;
  ln.NZ = nnz < 2 ? 2 : nnz;
  ln.span = 32;
  return true;
}

int run_prop(const kq_problem* p, bool backward, const double* pulses, const kq_c128* state0,
             kq_c128* stateT, kq_c128* store, void* stream, int k_lo = 0, int k_cnt = -1) {
  int rc = check_problem(p);
  if (rc) return rc;
  if (k_cnt < 0) k_cnt = p->K - k_lo;
  if (k_lo < 0 || k_cnt < 1 || k_lo + k_cnt > p->K)
    return fail(KQ_ERR_ARG, "invalid objective range [%d, %d) of %d", k_lo, k_lo + k_cnt, p->K);
  if (!pulses && p->L > 0) return fail(KQ_ERR_ARG, "pulses is NULL");
  if (!state0) return fail(KQ_ERR_ARG, "initial state is NULL");
  if (!stateT && !store) return fail(KQ_ERR_ARG, "both outputs are NULL");
  int dev;
  rc = device_init(&dev);
  if (rc) return rc;
  Plan pl;
  kq_problem sub = *p;   // launch geometry for the objectives actually propagated
  sub.K = k_cnt;
  rc = make_plan(&sub, false, false, g_dev[dev].sms, pl);
  if (rc) return rc;
  KqSweepArgs a = base_args(p);
  a.k_lo = k_lo;
  a.k_cnt = k_cnt;
  a.ops = reinterpret_cast<const cplx*>(backward ? p->ops_adj : p->ops);
  a.pulses = pulses;
  a.state0 = reinterpret_cast<const cplx*>(state0);
  a.stateT = reinterpret_cast<cplx*>(stateT);
  a.store = reinterpret_cast<cplx*>(store);
  a.backward = backward ? 1 : 0;
  const int fsel = p->is_super ? 2 : (backward ? 1 : 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pl.family == 2) {
    pl.grid = k_cnt;
    const size_t extra = csr_stage_bytes(p, false, pl.smem);
    pl.smem += extra;
    return kq_launch_csr(a, csr_of(p), pl, fsel, false, extra > 0, st);
  }
  // few objectives with small or sparse generators: entries-in-registers kernels
  // (kq_lanes.cuh), time-parallel over segments of the grid
  {
    KqLanes ln;
    if (!(pl.family == 0 && pl.spec) && pl.family != 2 && k_cnt <= 64 && prop_lanes_plan(p, ln)) {
      const int capacity = g_dev[dev].sms * KQ_LN_WMAX * ln.G;   // tasks of one wave
      int nseg = 1;
      if (!g_disable_segments && p->NT >= 64)
        nseg = std::min(128, std::min(p->NT / 8, capacity / (k_cnt * (p->N + 1))));
      auto warps_for = [&](int n_task) {
        return std::max(1, std::min(KQ_LN_WMAX, (n_task + ln.G - 1) / ln.G));
      };
      if (nseg < 4) {
        ln.W = warps_for(k_cnt);
        return kq_launch_prop_lanes(a, ln, fsel, k_cnt, st);
      }
      a.seg_len = (p->NT + nseg - 1) / nseg;
      nseg = (p->NT + a.seg_len - 1) / a.seg_len;
      a.seg_count = nseg;
      const size_t nP = (size_t)nseg * p->K * p->N * p->N, nB = (size_t)(nseg + 1) * p->K * p->N;
      void* scratch = nullptr;
      rc = get_scratch(dev, (nP + nB) * sizeof(cplx), &scratch);
      if (rc) return rc;
      a.seg_P = reinterpret_cast<cplx*>(scratch);
      a.seg_B = a.seg_P + nP;
      // pass 1: the N basis vectors through every segment -> segment propagators
      KqSweepArgs a1 = a;
      a1.seg_pass = 1;
      a1.store = nullptr;
      a1.stateT = nullptr;
      ln.W = warps_for(k_cnt * nseg * p->N);
      rc = kq_launch_prop_lanes(a1, ln, fsel, k_cnt * nseg * p->N, st);
      if (rc) return rc;
      {   // boundary states of the segments (sweep order)
        const int bt = round_up(p->N, 32);
        const size_t sm = (size_t)2 * p->N * sizeof(cplx);
        if (p->N <= 16)
          k_seg_chain_rows<16><<<k_cnt, bt, sm, st>>>(a, nseg);
        else
          k_seg_chain_rows<32><<<k_cnt, bt, sm, st>>>(a, nseg);
        KQ_CUDA(cudaGetLastError());
      }
      // pass 2: every segment from its boundary state, all states stored
      KqSweepArgs a2 = a;
      a2.seg_pass = 2;
      ln.W = warps_for(k_cnt * nseg);
      return kq_launch_prop_lanes(a2, ln, fsel, k_cnt * nseg, st);
    }
  }
  // few generic objectives (several terms): the thread-per-objective kernel would walk
  // through the grid with one thread; the lane-per-row family propagates segments of the
  // grid concurrently instead
  bool rows = false;
  if (pl.family == 0 && !pl.spec && g_lanes && p->N >= 2 && k_cnt * (p->N + 1) <= 64 &&
      p->NT >= 64 && !g_disable_segments) {
    rows = true;
    rc = make_plan(&sub, false, false, g_dev[dev].sms, pl, true);
    if (rc) return rc;
  }
  if (pl.family == 0) {
    if (!pl.spec) return kq_launch_prop_small(a, pl, fsel, st);
    // time-parallel propagation: segments of seg_len steps run concurrently
    // (pass 1: segment propagators, chain, pass 2: states), SURVEY.md §5.7
    // (with many objectives the sweep is bound by throughput, not by the chain: segments would
    // only multiply the work by N + 1)
    int nseg = 1;
    if (p->NT >= 64 && !g_disable_segments && k_cnt <= 16384) {
      a.seg_len = std::max(16, (p->NT + 63) / 64);
      nseg = (p->NT + a.seg_len - 1) / a.seg_len;
      const size_t nP = (size_t)nseg * p->K * p->N * p->N, nB = (size_t)(nseg + 1) * p->K * p->N;
      void* scratch = nullptr;
      rc = get_scratch(dev, (nP + nB) * sizeof(cplx), &scratch);
      if (rc) return rc;
      a.seg_P = reinterpret_cast<cplx*>(scratch);
      a.seg_B = a.seg_P + nP;
    }
    // real generator matrices in Hilbert space: purely imaginary f*A
    if (p->real_ops && !p->is_super) return kq_launch_prop_spec_re(a, pl, fsel, nseg, st);
    return kq_launch_prop_spec(a, pl, fsel, nseg, st);
  }
  // lane-per-row family: time-parallel propagation when the objectives alone
  // leave most of the GPU idle (one wave of tasks holds capacity objectives;
  // the segmented sweep does N + 1 times the work)
  Plan plmax;                    // the largest CTA the kernels take: tasks per wave
  sub.K = 1 << 20;
  rc = make_plan(&sub, false, false, g_dev[dev].sms, plmax, rows);
  if (rc) return rc;
  const int capacity = g_dev[dev].sms * (plmax.block / 32) * plmax.geom.G;
  int nseg = 1;
  if (!g_disable_segments && p->NT >= 64)
    nseg = std::min(128, std::min(p->NT / (rows ? 8 : 16), capacity / (k_cnt * (p->N + 1))));
  if (nseg < 4) return launch_warp(a, pl, fsel, false, false, st);
  a.seg_len = (p->NT + nseg - 1) / nseg;
  nseg = (p->NT + a.seg_len - 1) / a.seg_len;
  a.seg_count = nseg;
  {
    const size_t nP = (size_t)nseg * p->K * p->N * p->N, nB = (size_t)(nseg + 1) * p->K * p->N;
    void* scratch = nullptr;
    rc = get_scratch(dev, (nP + nB) * sizeof(cplx), &scratch);
    if (rc) return rc;
    a.seg_P = reinterpret_cast<cplx*>(scratch);
    a.seg_B = a.seg_P + nP;
  }
  // pass 1: the N basis vectors through every segment -> segment propagators
  KqSweepArgs a1 = a;
  a1.seg_pass = 1;
  a1.store = nullptr;
  a1.stateT = nullptr;
  Plan pl1;
  sub.K = k_cnt * nseg * p->N;
  rc = make_plan(&sub, false, false, g_dev[dev].sms, pl1, rows);
  if (rc) return rc;
  rc = launch_warp(a1, pl1, fsel, false, false, st);
  if (rc) return rc;
  // boundary states of the segments (sweep order)
  {
    const int bt = round_up(p->N, 32);
    const size_t sm = (size_t)2 * p->N * sizeof(cplx);
    if (p->N <= 16)
      k_seg_chain_rows<16><<<k_cnt, bt, sm, st>>>(a, nseg);
    else if (p->N <= 32)
      k_seg_chain_rows<32><<<k_cnt, bt, sm, st>>>(a, nseg);
    else
      k_seg_chain_rows<0><<<k_cnt, bt, sm, st>>>(a, nseg);
    KQ_CUDA(cudaGetLastError());
  }
  // pass 2: every segment from its boundary state, all states stored
  KqSweepArgs a2 = a;
  a2.seg_pass = 2;
  Plan pl2;
  sub.K = k_cnt * nseg;
  rc = make_plan(&sub, false, false, g_dev[dev].sms, pl2, rows);
  if (rc) return rc;
  return launch_warp(a2, pl2, fsel, false, false, st);
}

// lane = (objective, row), a few entries per row in registers: few objectives with small or
// sparse generators (kq_lanes.cuh); first order, one GPU.  kq_problem.row_nnz (0 = unknown:
// dense rows) is the largest number of non-zero columns in a row of the union pattern of an
// objective's terms, the diagonal included.
bool lanes_plan(const kq_problem* p, const KqSweepArgs& a, KqLanes& ln) {
  if (!g_lanes || a.world > 1 || p->N < 2 || p->N > 32 || p->L < 1 || p->L > KQ_LN_LMAX ||
      p->M > KQ_MMAX_SMALL || !p->ops || !p->mu)
    return false;
  const int nnz = p->row_nnz > 0 ? p->row_nnz : p->N;
  if (nnz > KQ_LN_NZMAX) return false;
  int NP = 2;
  while (NP < p->N) NP <<= 1;
  const int G = 32 / NP, W = (p->K + G - 1) / G;
  if (W > KQ_LN_WMAX) return false;
  ln.NP = NP;
  ln.G = G;
  ln.W = W;
  ln.NZ = nnz < 2 ? 2 : nnz;
  int span = NP;
  if (W > 1) {
    span = 32;
  } else {
    while (span < p->K * NP) span <<= 1;
  }
  ln.span = span;
  ln.zeta = nullptr;
  ln.scal = nullptr;
  return true;
}
int launch_lanes_update(const kq_problem* p, const KqSweepArgs& a, KqLanes& ln, int fsel,
                        cudaStream_t st) {
  int dev = 0;
  KQ_CUDA(cudaGetDevice(&dev));
  const size_t zb =
      ((size_t)p->NT * (p->L + 1) * ln.W * 32 * sizeof(cplx) + 255) / 256 * 256;
  void* base = nullptr;
  const int rc = get_scratch(dev, zb + (size_t)p->NT * KQ_LN_SC * sizeof(double), &base, 2);
  if (rc) return rc;
  ln.zeta = reinterpret_cast<cplx*>(base);
  ln.scal = reinterpret_cast<double*>(static_cast<char*>(base) + zb);
  return kq_launch_lanes(a, ln, fsel, st);
}

// The sequential update/forward sweep kernels (one time step after the other).
int launch_sequential_update(const kq_problem* p, const KqSweepArgs& a, const Plan& pl, int fsel,
                             bool second, cudaStream_t st) {
  if (pl.family == 2) {
    if (second) return fail(KQ_ERR_UNSUPPORTED, "second order is not built for N > 64");
    if (a.world > 1) return fail(KQ_ERR_UNSUPPORTED, "N > 64 is single-GPU");
    Plan pls = pl;
    const size_t extra = csr_stage_bytes(p, true, pl.smem);
    pls.smem += extra;
    return kq_launch_csr(a, csr_of(p), pls, fsel, true, extra > 0, st);
  }
  if (pl.family == 0) {
    if (!pl.spec) {
      KqLanes ln;
      if (!second && lanes_plan(p, a, ln)) return launch_lanes_update(p, a, ln, fsel, st);
      return kq_launch_fwupd_small(a, pl, fsel, second, st);
    }
    if (p->real_ops && !p->is_super) {
      // many two-level objectives (more than one CTA of the sequential kernel holds): the
      // kernel built around the grid-wide reduction (kq_sat.cuh)
      if (g_sat && p->N == 2 && !second && a.world == 1 && pl.grid > 1) {
        int dev = 0;
        KQ_CUDA(cudaGetDevice(&dev));
        if (g_dev[dev].coop && kq_sat_kpc(p->K, g_dev[dev].sms)) {
          KqSweepArgs b = a;
          b.pic_timing = g_picard_timing;
          return kq_launch_sat(b, g_dev[dev].sms, st);
        }
      }
      switch (p->N) {
        case 2: return kq_launch_fwupd_spec2_re(a, pl, fsel, second, st);
        case 3: return kq_launch_fwupd_spec3_re(a, pl, fsel, second, st);
        default: return kq_launch_fwupd_spec4_re(a, pl, fsel, second, st);
      }
    }
    switch (p->N) {
      case 2: return kq_launch_fwupd_spec2(a, pl, fsel, second, st);
      case 3: return kq_launch_fwupd_spec3(a, pl, fsel, second, st);
      default: return kq_launch_fwupd_spec4(a, pl, fsel, second, st);
    }
  }
  {
    // N >= 5 with sparse rows (the transmon of notebook 05: N = 17, three entries per row)
    KqLanes ln;
    if (!second && p->row_nnz > 0 && lanes_plan(p, a, ln))
      return launch_lanes_update(p, a, ln, fsel, st);
  }
  return launch_warp(a, pl, fsel, second, true, st);
}

// ---- boundary condition / overlaps --------------------------------------
__global__ void k_overlaps(int K, int N, const cplx* __restrict__ a, const cplx* __restrict__ b,
                           cplx* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  cplx acc = c_zero();
  for (int i = 0; i < N; ++i) acc = c_fma_conj(a[(size_t)k * N + i], b[(size_t)k * N + i], acc);
  out[k] = acc;
}

// Copy the column block [col0, col0+ncol) of every row of this rank's store to
// the same place in every peer's store (wide P2P stores over NVLink).
struct KqScatterArgs {
  cplx* dst[KQ_MAX_WORLD];
  int n_peer, self, rows, ncol;
  size_t row_stride, col0;
};
__global__ void k_scatter_rows(const KqScatterArgs a) {
  const long long total = (long long)a.rows * a.ncol;
  const cplx* src = a.dst[a.self];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const size_t idx = (size_t)(i / a.ncol) * a.row_stride + a.col0 + (size_t)(i % a.ncol);
    const cplx v = src[idx];
    for (int g = 0; g < a.n_peer; ++g)
      if (g != a.self) a.dst[g][idx] = v;
  }
}

// Cross-GPU barrier on the stream: every rank announces `tag` in every peer's
// exchange buffer and waits until all ranks have announced it in its own.
// Kernel boundaries order it after the P2P stores of the preceding kernel.
__global__ void k_comm_barrier(KqSlot* const* peer_slots, int rank, int world, uint32_t tag,
                               int* status) {
  const int r = threadIdx.x;
  const size_t off = (size_t)2 * KQ_MAX_WORLD * KQ_LMAX;   // after the per-step slots
  bool failed = false;
  if (r < world) {
    __threadfence_system();
    slot_store(peer_slots[r] + off + rank, 0.0, tag);
    slot_wait(peer_slots[rank] + off + r, tag, failed);
  }
  if (failed && status) atomicExch(status, (int)-4);
}

// chi_k(T) of functionals.py:177-197 (ss), 225-253 (sm), 293-317 (re),
// 389-437 (hs), then chi/||chi|| and ||chi|| (optimize.py:407-410).
__global__ void k_chi_boundary(int K, int N, int kind, int K_total, const cplx* __restrict__ phiT,
                               const cplx* __restrict__ targets, const cplx* __restrict__ tau,
                               const double* __restrict__ weights,
                               const cplx* __restrict__ tau_sum, cplx* __restrict__ chi,
                               double* __restrict__ norms) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double w = weights ? weights[k] : 1.0;
  cplx c;  // chi_k = c * target_k   (hs: c * (target_k - phi_k))
  if (kind == KQ_CHI_RE || kind == KQ_CHI_HS) {
    c = c_make(w * (1.0 / (2.0 * (double)K_total)), 0.0);
  } else if (kind == KQ_CHI_SS) {
    const cplx t = tau[k];
    c = c_make(t.x / (double)K_total * w, t.y / (double)K_total * w);
  } else {
    const double f = (1.0 / ((double)K_total * (double)K_total)) * w;
    const cplx s = tau_sum[0];
    c = c_make(f * s.x, f * s.y);
  }
  double nrm2 = 0.0;
  for (int i = 0; i < N; ++i) {
    cplx t = targets[(size_t)k * N + i];
    if (kind == KQ_CHI_HS) t = c_sub(t, phiT[(size_t)k * N + i]);
    const cplx v = c_make(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
    chi[(size_t)k * N + i] = v;
    nrm2 = fma(v.x, v.x, nrm2);
    nrm2 = fma(v.y, v.y, nrm2);
  }
  const double nrm = sqrt(nrm2);
  norms[k] = nrm;
  for (int i = 0; i < N; ++i) {
    cplx v = chi[(size_t)k * N + i];
    // a target reached exactly gives chi = 0: keep it zero instead of 0/0 (the reference
    // divides by zero there, optimize.py:410)
    chi[(size_t)k * N + i] = nrm > 0.0 ? c_make(v.x / nrm, v.y / nrm) : c_zero();
  }
}

// One Krotov iteration as a sequence of launches for problems the delta-polynomial family
// serves (few objectives, one control, first order): plan (chi boundary) | build | segment
// products | expand | sweep | epilogue (tau).  A declined or failed iteration is reported
// like a fixed point that did not converge; the caller repeats it with the sweep calls.
int composite_iteration(const kq_problem* p, DpPlan& dp, KqSweepArgs a, int chi_kind,
                        void* workspace, int dev, cudaStream_t st) {
  DpScratch s;
  int rc = dpoly_scratch(p, dp, dev, s);
  if (rc) return rc;
  dp.d.chain = 1;
  dp.d.segP = s.segP;
  dp.d.X = a.Xout ? a.Xout : s.X;
  dp.d.tsum = s.tsum;
  if (chi_kind >= 0) {
    dp.d.chi = a.chi_out ? a.chi_out : s.chi;
    dp.d.norms = a.chi_norms_out ? a.chi_norms_out : s.norms;
  } else {
    dp.d.chi = const_cast<cplx*>(a.chiT);
    dp.d.norms = const_cast<double*>(a.chi_norms);
  }
  if (!a.stateT) a.stateT = s.phiT;
  a.X = dp.d.X;
  a.chi_norms = dp.d.norms;
  a.pic_bw = 0;
  rc = launch_dpoly(p, a, dp, workspace, st);
  if (rc) return rc;
  return kq_launch_dpoly_epilogue(a, dp.d, 0, st);
}


// sum_j w_j tau_j (chis_sm, functionals.py:225-253), fixed order
__global__ void k_tau_sum(int K, const cplx* __restrict__ tau, const double* __restrict__ weights,
                          cplx* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double x = 0.0, y = 0.0;
  for (int j = 0; j < K; ++j) {
    const double w = weights ? weights[j] : 1.0;
    x = fma(w, tau[j].x, x);
    y = fma(w, tau[j].y, y);
  }
  out[0] = c_make(x, y);
}
// status words of a composite iteration that cannot decline: {status, 0, 0, 0}
__global__ void k_diag_ok(const int* status, int* diag_out) {
  if (threadIdx.x == 0) {
    diag_out[0] = *reinterpret_cast<const volatile int*>(status);
    diag_out[1] = 0;
    diag_out[2] = 0;
    diag_out[3] = 0;
  }
}

// One Krotov iteration as a sequence of launches for the entries-in-registers family
// (kq_lanes.cuh: few objectives, several controls or sparse rows, first order, one GPU):
// chi boundary | time-parallel backward sweep | pre-pass | update chain | tau.
int rows_iteration(const kq_problem* p, KqLanes& ln, KqSweepArgs a, int chi_kind, int K_total,
                   const cplx* tau_sum_in, const double* guess_pulses, void* workspace, int dev,
                   cudaStream_t st) {
  const int K = p->K, N = p->N, NT = p->NT;
  const size_t xb = ((size_t)(NT + 1) * K * N * sizeof(cplx) + 255) / 256 * 256;
  const size_t sb = ((size_t)K * N * sizeof(cplx) + 255) / 256 * 256;
  const size_t nb = ((size_t)K * sizeof(double) + 255) / 256 * 256;
  void* base = nullptr;
  int rc = get_scratch(dev, xb + 2 * sb + nb + 256, &base, 1);
  if (rc) return rc;
  char* cur = static_cast<char*>(base);
  cplx* sX = reinterpret_cast<cplx*>(cur);
  cur += xb;
  cplx* schi = reinterpret_cast<cplx*>(cur);
  cur += sb;
  cplx* sphiT = reinterpret_cast<cplx*>(cur);
  cur += sb;
  double* snorms = reinterpret_cast<double*>(cur);
  cur += nb;
  cplx* stsum = reinterpret_cast<cplx*>(cur);
  cplx* X = a.Xout ? a.Xout : sX;
  const cplx* chi = a.chiT;
  const double* norms = a.chi_norms;
  if (chi_kind >= 0) {
    cplx* chi_w = a.chi_out ? a.chi_out : schi;
    double* norms_w = a.chi_norms_out ? a.chi_norms_out : snorms;
    const cplx* tsum = tau_sum_in;
    if (chi_kind == KQ_CHI_SM && !tsum) {
      k_tau_sum<<<1, 32, 0, st>>>(K, a.tau_in, a.weights, stsum);
      KQ_CUDA(cudaGetLastError());
      tsum = stsum;
    }
    const int bt = 128;
    k_chi_boundary<<<(K + bt - 1) / bt, bt, 0, st>>>(K, N, chi_kind, K_total, a.phiT_in, a.targets,
                                                     a.tau_in, a.weights, tsum, chi_w, norms_w);
    KQ_CUDA(cudaGetLastError());
    chi = chi_w;
    norms = norms_w;
  }
  rc = run_prop(p, true, guess_pulses, reinterpret_cast<const kq_c128*>(chi), nullptr,
                reinterpret_cast<kq_c128*>(X), st);
  if (rc) return rc;
  if (!a.stateT) a.stateT = sphiT;
  a.X = X;
  a.chi_norms = norms;
  a.status = reinterpret_cast<int*>(workspace);
  a.cond_epoch = 0;
  rc = launch_lanes_update(p, a, ln, p->is_super ? 2 : 0, st);
  if (rc) return rc;
  if (a.tau_out && a.targets) {
    const int bt = 128;
    k_overlaps<<<(K + bt - 1) / bt, bt, 0, st>>>(K, N, a.targets, a.stateT, a.tau_out);
    KQ_CUDA(cudaGetLastError());
  }
  if (a.diag_out) {
    k_diag_ok<<<1, 32, 0, st>>>(a.status, a.diag_out);
    KQ_CUDA(cudaGetLastError());
  }
  return KQ_OK;
}

}  // namespace

extern "C" {

int kq_version(void) { return 104; }

int kq_set_option(const char* name, int value) {
  if (name && std::strcmp(name, "time_parallel") == 0) {
    g_disable_segments = (value == 0);
    return KQ_OK;
  }
  if (name && std::strcmp(name, "picard") == 0) {
    if (value < 0 || value > 2) return fail(KQ_ERR_ARG, "picard must be 0, 1 or 2");
    g_picard = value;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "picard_timing") == 0) {
    g_picard_timing = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "cooperative_launch") == 0) {
    g_kq_coop_launch = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "programmatic_launch") == 0) {
    g_kq_pdl_launch = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "dpoly") == 0) {
    if (value < 0 || value > 2) return fail(KQ_ERR_ARG, "dpoly must be 0, 1 or 2");
    g_dpoly = value;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "dpoly_debug") == 0) {
    g_dpoly_debug = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "lanes") == 0) {
    g_lanes = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "sat_min_kpc") == 0) {
    if (value < 2 || value > 896) return fail(KQ_ERR_ARG, "sat_min_kpc must be in [2, 896]");
    g_kq_sat_min_kpc = value;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "sat") == 0) {
    g_sat = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "small_rows") == 0) {
    g_small_rows = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "picard_history") == 0) {
    g_picard_history = value ? 1 : 0;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "picard_rtol_e15") == 0) {
    if (value < 1 || value > 1000000) return fail(KQ_ERR_ARG, "picard_rtol_e15 out of range");
    g_picard_rtol_e15 = value;
    return KQ_OK;
  }
  if (name && std::strcmp(name, "picard_maxit") == 0) {
    if (value < 1 || value > kPicMaxItCap) return fail(KQ_ERR_ARG, "picard_maxit out of range");
    g_picard_maxit = value;
    return KQ_OK;
  }
  return fail(KQ_ERR_ARG, "unknown option '%s'", name ? name : "(null)");
}

// ---- cross-GPU exchange buffers (CUDA IPC) --------------------------------
int kq_comm_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return fail(KQ_ERR_ARG, "invalid argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  KQ_CUDA(cudaMalloc(&p, bytes));
  KQ_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(KQ_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  std::memcpy(handle64, &h, 64);
  *ptr = p;
  return KQ_OK;
}

int kq_comm_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(KQ_ERR_ARG, "invalid argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  KQ_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KQ_OK;
}

int kq_comm_close(void* ptr) {
  if (ptr) KQ_CUDA(cudaIpcCloseMemHandle(ptr));
  return KQ_OK;
}

int kq_comm_free(void* ptr) {
  if (ptr) KQ_CUDA(cudaFree(ptr));
  return KQ_OK;
}

const char* kq_last_error(void) { return g_err.c_str(); }

size_t kq_comm_slot_bytes(const kq_problem* p) {
  size_t slots = xg_offset_slots();
  // time-parallel iteration with sharded objectives: [4][cta][rank][slice] sums
  if (p && p->NT > 0) slots += (size_t)4 * kMaxWorld * pic_stride(p);
  return slots * sizeof(KqSlot);
}

int kq_comm_barrier(const kq_comm* comm, uint32_t tag, void* workspace, void* stream) {
  if (!comm || comm->world < 1 || comm->world > kMaxWorld || !comm->slots)
    return fail(KQ_ERR_ARG, "invalid kq_comm");
  if (tag == 0) return fail(KQ_ERR_ARG, "barrier tag must be non-zero");
  k_comm_barrier<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<KqSlot* const*>(comm->slots), comm->rank, comm->world, tag,
      reinterpret_cast<int*>(workspace));
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}

size_t kq_workspace_bytes(const kq_problem* p) {
  // ... | header of the delta-polynomial iteration | its anchor pulse and step records
  return dpoly_header_offset(p) + 128 + dpoly_workspace_bytes(p);
}

size_t kq_dpoly_header_offset(const kq_problem* p) { return dpoly_header_offset(p); }

int kq_plan(const kq_problem* p, int32_t* family, int32_t* grid, int32_t* block,
            int32_t* smem_bytes) {
  int rc = check_problem(p);
  if (rc) return rc;
  Plan pl;
  rc = make_plan(p, true, false, 148, pl);
  if (rc) return rc;
  {
    // the families that stand in front of the generic kernels (launch_sequential_update)
    KqSweepArgs probe = base_args(p);
    KqLanes ln;
    const bool generic = pl.family == 1 || (pl.family == 0 && !pl.spec);
    if (generic && (pl.family == 0 || p->row_nnz > 0) && lanes_plan(p, probe, ln)) {
      pl.family = 3;
      pl.grid = 1;
      pl.block = ln.W * 32;
      pl.smem = 2 * KQ_LN_LMAX * 8 * sizeof(double);
    } else if (pl.family == 0 && pl.spec && g_sat && p->real_ops && !p->is_super && p->N == 2 &&
               pl.grid > 1 && kq_sat_kpc(p->K, 148)) {
      const int kpc = kq_sat_kpc(p->K, 148);
      pl.family = 4;
      pl.grid = (p->K + kpc - 1) / kpc;
      pl.block = 512;
      pl.smem = kq_sat_smem(kpc);
    }
  }
  if (family) *family = pl.family;
  if (grid) *grid = pl.grid;
  if (block) *block = pl.block;
  if (smem_bytes) *smem_bytes = (int32_t)pl.smem;
  return KQ_OK;
}

int kq_plan_fused(const kq_problem* p, int32_t* grid, int32_t* block, int32_t* chunk,
                  int32_t* smem_bytes) {
  if (!p || p->K < 1 || p->N < 1 || p->NT < 1) return fail(KQ_ERR_ARG, "invalid problem sizes");
  PicPlan pp;
  if (!picard_plan(p, 148, pp))
    return fail(KQ_ERR_UNSUPPORTED, "problem is outside the time-parallel kernel family");
  if (grid) *grid = pp.grid;
  if (block) *block = pp.block;
  if (chunk) *chunk = pp.W;
  if (smem_bytes) *smem_bytes = (int32_t)pp.smem;
  return KQ_OK;
}

int kq_propagate_forward(const kq_problem* p, const double* pulses, const kq_c128* state0,
                         kq_c128* stateT, kq_c128* store, void* stream) {
  return run_prop(p, false, pulses, state0, stateT, store, stream);
}

int kq_sweep_backward(const kq_problem* p, const double* guess_pulses, const kq_c128* chiT,
                      kq_c128* X, void* stream) {
  if (!X) return fail(KQ_ERR_ARG, "X is NULL");
  return run_prop(p, true, guess_pulses, chiT, nullptr, X, stream);
}

int kq_sweep_backward_range(const kq_problem* p, const double* guess_pulses,
                            const kq_c128* chiT, void* const* X_peers, int32_t n_peer,
                            int32_t self, int32_t k_lo, int32_t k_cnt, void* stream) {
  if (!X_peers || n_peer < 1 || n_peer > kMaxWorld || self < 0 || self >= n_peer)
    return fail(KQ_ERR_ARG, "invalid peer table");
  int rc = run_prop(p, true, guess_pulses, chiT, nullptr,
                    reinterpret_cast<kq_c128*>(X_peers[self]), stream, k_lo, k_cnt);
  if (rc || n_peer == 1) return rc;
  // broadcast this rank's block of columns of X to every peer (P2P stores)
  KqScatterArgs sa;
  sa.n_peer = n_peer;
  sa.self = self;
  for (int g = 0; g < n_peer; ++g) sa.dst[g] = reinterpret_cast<cplx*>(X_peers[g]);
  sa.rows = p->NT + 1;
  sa.row_stride = (size_t)p->K * p->N;
  sa.col0 = (size_t)k_lo * p->N;
  sa.ncol = k_cnt * p->N;
  const long long total = (long long)sa.rows * sa.ncol;
  int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
  k_scatter_rows<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(sa);
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}

int kq_sweep_forward_update(const kq_problem* p, const double* guess_pulses, double* opt_pulses,
                            const kq_c128* X, const double* chi_norms, const kq_c128* phi0,
                            kq_c128* phiT, const double* sigma, const kq_c128* Phi0,
                            kq_c128* Phi1, double* g_a, const kq_comm* comm, void* workspace,
                            uint32_t epoch, void* stream) {
  int rc = check_problem(p);
  if (rc) return rc;
  if (!guess_pulses || !opt_pulses || !X || !chi_norms || !phi0 || !g_a || !workspace)
    return fail(KQ_ERR_ARG, "NULL argument to kq_sweep_forward_update");
  if ((!p->mu && !p->sparse) || !p->shape || !p->lambda_a)
    return fail(KQ_ERR_ARG, "problem lacks mu/shape/lambda_a");
  if (p->L < 1) return fail(KQ_ERR_ARG, "no pulses to update");
  const bool second = sigma != nullptr;
  if (second && !Phi0) return fail(KQ_ERR_ARG, "second order needs Phi0");
  int dev;
  rc = device_init(&dev);
  if (rc) return rc;
  Plan pl;
  rc = make_plan(p, true, second, g_dev[dev].sms, pl);
  if (rc) return rc;
  if (pl.grid > kMaxBlocks) return fail(KQ_ERR_UNSUPPORTED, "too many CTAs (%d)", pl.grid);
  KqSweepArgs a = base_args(p);
  a.ops = reinterpret_cast<const cplx*>(p->ops);
  a.pulses = guess_pulses;
  a.opt_pulses = opt_pulses;
  a.state0 = reinterpret_cast<const cplx*>(phi0);
  a.stateT = reinterpret_cast<cplx*>(phiT);
  a.store = second ? reinterpret_cast<cplx*>(Phi1) : nullptr;
  a.X = reinterpret_cast<const cplx*>(X);
  a.chi_norms = chi_norms;
  a.sigma = sigma;
  a.Phi0 = reinterpret_cast<const cplx*>(Phi0);
  a.g_a = g_a;
  a.status = reinterpret_cast<int*>(workspace);
  a.slots = reinterpret_cast<KqSlot*>(static_cast<char*>(workspace) + kStatusBytes);
  a.tag_base = epoch * (uint32_t)(p->NT + 1);
  if (comm && comm->world > 1) {
    if (comm->world > kMaxWorld || !comm->slots)
      return fail(KQ_ERR_ARG, "invalid kq_comm (world=%d)", comm->world);
    a.rank = comm->rank;
    a.world = comm->world;
    a.peer_slots = reinterpret_cast<KqSlot* const*>(comm->slots);
  }
  const int fsel = p->is_super ? 2 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  a.epoch = epoch ? epoch : 1u;
  // few objectives, one control: delta-polynomial sweep (kq_dpoly.cuh) with the sequential
  // kernel queued behind it as an in-stream conditional fall-back
  DpPlan dp;
  const bool use_dpoly = !second && a.world == 1 && dpoly_preferred(p) && dpoly_plan(p, dp);
  if (use_dpoly) {
    // backward states from the caller's X; chi norms as given
    DpScratch s;
    rc = dpoly_scratch(p, dp, dev, s);
    if (rc) return rc;
    dp.d.chain = 0;
    dp.d.segP = s.segP;
    dp.d.X = const_cast<cplx*>(a.X);
    dp.d.chi = nullptr;
    dp.d.norms = const_cast<double*>(chi_norms);
    dp.d.tsum = s.tsum;
    a.chi_kind = -1;
    rc = launch_dpoly(p, a, dp, workspace, st);
    if (rc) return rc;
    a.cond_epoch = a.epoch;
  }
  // time-parallel fused sweep (kq_picard.cuh), with the sequential kernel
  // queued behind it as a conditional fall-back.  The sweep may be cut into
  // time windows that are solved one after the other (each a fixed-point
  // problem of its own, started from the final states of the window before):
  // needed when the state stores of the whole grid exceed shared memory, and
  // faster for strongly coupled problems, where the number of rounds grows with
  // the length of the interval (kq_problem.reserved = number of windows, 0 = as
  // few as fit).
  if (!use_dpoly && g_picard && pl.family == 0 && pl.spec && a.world == 1 && g_dev[dev].coop) {
    int nwin = p->reserved > 0 ? std::min(p->reserved, kPicMaxWindows) : 1;
    nwin = std::min(nwin, std::max(1, p->NT / 32));
    kq_problem pw = *p;
    PicPlan pp;
    bool ok = false;
    for (; nwin <= kPicMaxWindows && nwin <= std::max(1, p->NT / 8); nwin *= 2) {
      pw.NT = (p->NT + nwin - 1) / nwin;
      if (picard_plan(&pw, g_dev[dev].sms, pp)) {
        ok = true;
        break;
      }
      if (p->N < 2 || p->N > 4 || p->M != 2 || p->L != 1) break;   // not a size problem
    }
    if (ok && nwin > 1 && !phiT) ok = false;   // windows hand their final states on through phiT
    if (ok) {
      const int len = pw.NT;
      const size_t row = (size_t)p->K * p->N;
      for (int w = 0, n0 = 0; n0 < p->NT; ++w, n0 += len) {
        const int n1 = std::min(p->NT, n0 + len);
        pw.NT = n1 - n0;
        pw.dt = p->dt + n0;
        pw.shape = p->shape + n0;
        PicPlan ppw;
        if (!picard_plan(&pw, g_dev[dev].sms, ppw)) return fail(KQ_ERR_UNSUPPORTED, "window plan");
        KqSweepArgs b = a;
        b.NT = pw.NT;
        b.dt = pw.dt;
        b.shape = pw.shape;
        b.pulses = guess_pulses + n0;
        b.opt_pulses = opt_pulses + n0;
        b.X = a.X + (size_t)n0 * row;
        if (a.sigma) b.sigma = a.sigma + n0;
        if (a.Phi0) b.Phi0 = a.Phi0 + (size_t)n0 * row;
        if (a.store) b.store = a.store + (size_t)n0 * row;
        if (w > 0) b.state0 = reinterpret_cast<const cplx*>(phiT);
        b.pic_bw = 0;
        b.pic_accumulate = (w > 0) ? 1 : 0;
        rc = launch_picard(&pw, b, ppw, workspace, epoch, second, st, w);
        if (rc) return rc;
      }
      a.cond_epoch = epoch ? epoch : 1u;   // the sequential kernel below runs only on request
    }
  }
  if (!(use_dpoly && g_dpoly_debug)) {
    rc = launch_sequential_update(p, a, pl, fsel, second, st);
    if (rc) return rc;
  }
  if (use_dpoly) return kq_launch_dpoly_epilogue(a, dp.d, 1, st);
  return KQ_OK;
}

int kq_krotov_iteration(const kq_problem* p, int chi_kind, int32_t K_total,
                        const kq_c128* targets, const double* weights, const kq_c128* chiT,
                        const double* chi_norms, const kq_c128* tau_in, const kq_c128* phiT_in,
                        const double* guess_pulses, const double* prev_guess_pulses,
                        double* opt_pulses, const kq_c128* phi0, kq_c128* phiT_out, kq_c128* tau_out, kq_c128* X, kq_c128* chi_out,
                        double* chi_norms_out, const double* sigma, const kq_c128* Phi0,
                        kq_c128* Phi1, double* g_a, int32_t* diag_out, const kq_comm* comm,
                        const kq_c128* tau_sum, void* workspace, uint32_t epoch,
                        void* stream) {
  int rc = check_problem(p);
  if (rc) return rc;
  if (!guess_pulses || !opt_pulses || !phi0 || !g_a || !workspace)
    return fail(KQ_ERR_ARG, "NULL argument to kq_krotov_iteration");
  if ((!p->mu && !p->sparse) || !p->shape || !p->lambda_a)
    return fail(KQ_ERR_ARG, "problem lacks mu/shape/lambda_a");
  if (chi_kind < -1 || chi_kind > KQ_CHI_HS) return fail(KQ_ERR_ARG, "unknown chi kind %d", chi_kind);
  if (chi_kind < 0 && (!chiT || !chi_norms)) return fail(KQ_ERR_ARG, "chiT/chi_norms are NULL");
  if (chi_kind >= 0 && !targets) return fail(KQ_ERR_ARG, "built-in chi needs targets");
  if ((chi_kind == KQ_CHI_SS || chi_kind == KQ_CHI_SM) && !tau_in)
    return fail(KQ_ERR_ARG, "chis_ss/chis_sm need tau_in");
  if (chi_kind == KQ_CHI_HS && !phiT_in) return fail(KQ_ERR_ARG, "chis_hs needs phiT_in");
  const int world = (comm && comm->world > 1) ? comm->world : 1;
  if (world > 1 && (comm->world > kMaxWorld || !comm->slots || comm->rank < 0 ||
                    comm->rank >= comm->world))
    return fail(KQ_ERR_ARG, "invalid kq_comm (world=%d)", comm->world);
  if (world == 1 && chi_kind >= 0 && K_total != p->K)
    return fail(KQ_ERR_ARG, "K_total must equal K without a kq_comm");
  if (world > 1 && (K_total < p->K || (K_total + world - 1) / world < p->K))
    return fail(KQ_ERR_ARG, "K=%d objectives on this rank exceed ceil(K_total/world) = %d", p->K,
                (K_total + world - 1) / world);
  if (world > 1 && chi_kind == KQ_CHI_SM && !tau_sum)
    return fail(KQ_ERR_ARG, "chis_sm with sharded objectives needs tau_sum");
  const bool second = sigma != nullptr;
  if (second && !Phi0) return fail(KQ_ERR_ARG, "second order needs Phi0");
  int dev;
  rc = device_init(&dev);
  if (rc) return rc;
  PicPlan pp;
  // all ranks use the geometry of the largest block of objectives
  const int K_plan = world > 1 ? (K_total + world - 1) / world : p->K;
  DpPlan dp;
  const bool composite = !second && world == 1 && dpoly_preferred(p) && dpoly_plan(p, dp);
  const bool fixed_point =
      !composite && g_picard && g_dev[dev].coop && picard_plan(p, g_dev[dev].sms, pp, K_plan);
  // what neither takes: few objectives with several controls or sparse rows (kq_lanes.cuh)
  KqLanes ln;
  bool rows = false;
  // (not the M = 2, L = 1, N <= 4 family: what the one-launch kernel declines there -- long
  // grids -- is served better by the windowed fixed point of kq_sweep_forward_update)
  if (!composite && !fixed_point && !second && world == 1 &&
      !(p->N <= 4 && p->M == 2 && p->L == 1)) {
    KqSweepArgs probe = base_args(p);
    KqLanes lp;
    rows = p->ops && p->ops_adj && lanes_plan(p, probe, ln) && prop_lanes_plan(p, lp);
  }
  if (!composite && !fixed_point && !rows)
    return fail(KQ_ERR_UNSUPPORTED, "problem is outside the time-parallel kernel family");
  KqSweepArgs a = base_args(p);
  if (world > 1) {
    a.rank = comm->rank;
    a.world = world;
    a.peer_slots = reinterpret_cast<KqSlot* const*>(comm->slots);
    a.tau_sum = reinterpret_cast<const cplx*>(tau_sum);
  }
  a.ops = reinterpret_cast<const cplx*>(p->ops);
  a.ops_adj = reinterpret_cast<const cplx*>(p->ops_adj);
  a.pulses = guess_pulses;
  a.opt_pulses = opt_pulses;
  a.state0 = reinterpret_cast<const cplx*>(phi0);
  a.stateT = reinterpret_cast<cplx*>(phiT_out);
  a.store = second ? reinterpret_cast<cplx*>(Phi1) : nullptr;
  a.chi_norms = chi_norms;
  a.sigma = sigma;
  a.Phi0 = reinterpret_cast<const cplx*>(Phi0);
  a.g_a = g_a;
  a.pic_bw = 1;
  a.pic_hint = prev_guess_pulses;
  a.chi_kind = chi_kind;
  a.K_total = K_total;
  a.chiT = reinterpret_cast<const cplx*>(chiT);
  a.targets = reinterpret_cast<const cplx*>(targets);
  a.weights = weights;
  a.tau_in = reinterpret_cast<const cplx*>(tau_in);
  a.tau_out = reinterpret_cast<cplx*>(tau_out);
  a.phiT_in = reinterpret_cast<const cplx*>(phiT_in);
  a.Xout = reinterpret_cast<cplx*>(X);
  a.chi_out = reinterpret_cast<cplx*>(chi_out);
  a.chi_norms_out = chi_norms_out;
  a.diag_out = diag_out;
  if (composite) {
    a.epoch = epoch ? epoch : 1u;
    a.status = reinterpret_cast<int*>(workspace);
    return composite_iteration(p, dp, a, chi_kind, workspace, dev,
                               static_cast<cudaStream_t>(stream));
  }
  if (rows) {
    a.epoch = epoch ? epoch : 1u;
    return rows_iteration(p, ln, a, chi_kind, K_total, reinterpret_cast<const cplx*>(tau_sum),
                          guess_pulses, workspace, dev, static_cast<cudaStream_t>(stream));
  }
  return launch_picard(p, a, pp, workspace, epoch, second, static_cast<cudaStream_t>(stream));
}

int kq_chi_boundary(const kq_problem* p, int kind, int32_t K_total, const kq_c128* phiT,
                    const kq_c128* targets, const kq_c128* tau, const double* weights,
                    const kq_c128* tau_sum, kq_c128* chi_out, double* chi_norms, void* stream) {
  if (!p || !targets || !chi_out || !chi_norms) return fail(KQ_ERR_ARG, "NULL argument");
  if (kind < KQ_CHI_RE || kind > KQ_CHI_HS) return fail(KQ_ERR_ARG, "unknown chi kind %d", kind);
  if (kind == KQ_CHI_SS && !tau) return fail(KQ_ERR_ARG, "chis_ss needs tau");
  if (kind == KQ_CHI_SM && !tau_sum) return fail(KQ_ERR_ARG, "chis_sm needs tau_sum");
  if (kind == KQ_CHI_HS && !phiT) return fail(KQ_ERR_ARG, "chis_hs needs phiT");
  if (K_total < p->K) return fail(KQ_ERR_ARG, "K_total < K");
  const int bt = 128;
  k_chi_boundary<<<(p->K + bt - 1) / bt, bt, 0, static_cast<cudaStream_t>(stream)>>>(
      p->K, p->N, kind, K_total, reinterpret_cast<const cplx*>(phiT),
      reinterpret_cast<const cplx*>(targets), reinterpret_cast<const cplx*>(tau), weights,
      reinterpret_cast<const cplx*>(tau_sum), reinterpret_cast<cplx*>(chi_out), chi_norms);
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}

int kq_fetch_results(void* dst_host, const void* src_device, size_t nbytes, void* stream) {
  if (!dst_host || !src_device) return fail(KQ_ERR_ARG, "NULL argument to kq_fetch_results");
  KQ_CUDA(cudaMemcpyAsync(dst_host, src_device, nbytes, cudaMemcpyDeviceToHost,
                          static_cast<cudaStream_t>(stream)));
  return KQ_OK;
}

int kq_overlaps(int32_t K, int32_t N, const kq_c128* a, const kq_c128* b, kq_c128* out,
                void* stream) {
  if (K < 1 || N < 1 || !a || !b || !out) return fail(KQ_ERR_ARG, "invalid argument");
  const int bt = 128;
  k_overlaps<<<(K + bt - 1) / bt, bt, 0, static_cast<cudaStream_t>(stream)>>>(
      K, N, reinterpret_cast<const cplx*>(a), reinterpret_cast<const cplx*>(b),
      reinterpret_cast<cplx*>(out));
  KQ_CUDA(cudaGetLastError());
  return KQ_OK;
}

}  // extern "C"
