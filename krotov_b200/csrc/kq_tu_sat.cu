// Update sweep for many two-level objectives with a real generator (kq_sat.cuh): launch.
#include "kq_host.cuh"
#include "kq_sat.cuh"

KQ_DEFINE_TABLES_UPLOAD(kq_tables_upload_sat)

size_t kq_sat_smem(int kpc) {
  return 56 * sizeof(double) + 2 * KQ_SAT_RING * sizeof(uint64_t) + 4 * KQ_NTC * sizeof(double) + KQ_NTC +
         (size_t)KQ_SAT_RING * kpc * 2 * sizeof(cplx) + (size_t)(kpc + 1) * 4 * sizeof(double);
}

// objectives per CTA: at most KQ_SAT_BT * KQ_SAT_OPT; 0 if the problem does not fit `sms` CTAs
int g_kq_sat_min_kpc = 64;   // kq_set_option("sat_min_kpc", v): fewer, fuller CTAs (experiments)
int kq_sat_kpc(int K, int sms) {
  int kpc = std::max(g_kq_sat_min_kpc, (K + sms - 1) / sms);
  kpc = (kpc + 1) & ~1;
  // mailboxes [2][grid][grid] must fit the workspace's slot area (2 * 4096 * KQ_LMAX slots)
  const long long grid = (K + kpc - 1) / kpc;
  if (2 * grid * grid > 2LL * KQ_MAX_BLOCKS * KQ_LMAX || grid > 32 * KQ_SAT_NG) return 0;
  return kpc <= KQ_SAT_BT * KQ_SAT_OPT ? kpc : 0;
}

int kq_launch_sat(const KqSweepArgs& a, int sms, cudaStream_t st) {
  int kpc = kq_sat_kpc(a.K, sms);
  if (!kpc) return kq_fail(KQ_ERR_UNSUPPORTED, "K=%d exceeds one co-resident grid", a.K);
  KqPlan pl = {};
  pl.grid = (a.K + kpc - 1) / kpc;
  pl.block = KQ_SAT_THREADS;
  pl.smem = kq_sat_smem(kpc);
  void* params[] = {(void*)&a, (void*)&kpc};
  return launch(k_fwupd_sat, pl, pl.grid > 1, st, params);
}
