// Update / forward sweep (optimize.py:449-500 of the reference) for MANY two-level objectives
// with a real generator (the detuning ensemble of BASELINE configs[3] scaled up: K = 131 072
// objectives, backward-state store 4.2 GB): the regime where the sweep streams from HBM.
//
// Every time step needs the sum over ALL objectives before any of them can move on
// (optimize.py:454-470), so a step costs one grid-wide all-reduce whatever else is done; the
// kernel is built around making that reduction short and keeping everything else off it:
//   * one CTA per SM (co-resident grid); warps 2..15 are CONSUMERS with TWO objectives per
//     thread (14 warps instead of 32 at the CTA barriers of a step, 128 registers per thread
//     -- the 1024-thread kernel of kq_spec.cuh spills at 64 --, two independent chains per
//     thread); warp 0 owns no objectives: it reduces the warps' partial sums and publishes the
//     CTA's sum (objectives in the exchange warp would put their work on the critical path);
//     warp 1 issues the TMA copies (measured: a bulk-copy issue
//     holds its warp, and the warp's next shared-memory access, for 500-3000 cycles -- from the
//     exchange warp that delay lands on every CTA of the grid);
//   * the backward states of a time step are one contiguous row per CTA (time-major
//     [nt][K][N] layout) streamed HBM -> shared memory by TMA bulk copies, KQ_SAT_RING steps
//     ahead of their use; eta = mu^dag chi ||chi|| of the next step is formed while the
//     reduction of this one is under way;
//   * CTAs exchange their partial sums through flag-tagged 16-byte slots in L2 (kq_common.cuh):
//     every CTA PUSHES its sum into a mailbox per CTA (two 64-bit max-reductions performed at
//     L2: {tag, half of the value}; plain stores took 0.6k cycles longer to become visible), so
//     that a mailbox's lines are polled by one CTA only; FIVE consumer warps -- idle while the
//     exchange is in flight -- poll 32 slots each with ONE load per lane (a coherent load costs
//     ~1000 cycles here and the loads of one thread complete one after the other: five slots
//     per lane in one warp cost 1.6k cycles per poll) and every CTA adds the values in the same
//     order: identical pulses in every CTA, no floating-point atomics.  Measured and
//     rejected: several polling loads in flight per lane, polling with atom.or, a delay in
//     front of the first poll, fewer and fuller CTAs (the hop costs the same for 19 and for
//     148 participants: it is latency, not contention);
//   * the step itself is the closed form of exp(iR) for a real 2 x 2 generator (kq_spec.cuh),
//     its series degree planned from the guess pulse and verified after the update; B200
//     issues 58 DFMA per clock and SM, so the loop body is kept to about 50 FP64 operations
//     per objective and step (mu is real for these problems: kq_problem.real_ops).
#pragma once
#include "kq_spec.cuh"

#define KQ_SAT_BT 448       // consumer threads (warps 2..15); warp 0: exchange, warp 1: TMA
#define KQ_SAT_THREADS (KQ_SAT_BT + 64)
#define KQ_SAT_OPT 2
#define KQ_SAT_RING 4
#define KQ_SAT_NG 5       // consumer warps that gather the mailbox: 32 slots each (grid <= 160)

// Publication of a slot as two 64-bit REDUCTIONS (max) performed at L2: {tag, half of the
// value} with the tag in the high word -- tags only grow, so the maximum is the newest entry.
// A reduction is carried out where the pollers read; plain stores were measured to take
// 1.5k+ cycles until a poller on another SM sees them.
__device__ __forceinline__ void sat_slot_store(KqSlot* p, double v, uint32_t tag) {
  const unsigned long long w0 = ((unsigned long long)tag << 32) | (uint32_t)__double2loint(v);
  const unsigned long long w1 = ((unsigned long long)tag << 32) | (uint32_t)__double2hiint(v);
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(w0) : "memory");
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(reinterpret_cast<char*>(p) + 8), "l"(w1)
               : "memory");
}
__device__ __forceinline__ void sat_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier A / staging barrier: consumers + exchange warp (the producer warp runs on its own)
__device__ __forceinline__ void sat_bar_a() {
  asm volatile("bar.sync 3, %0;" ::"n"(KQ_SAT_BT + 32) : "memory");
}

// One objective: R = h T0f + heps T1f (real 2 x 2, column-major [[a, b], [c, d]]) enters the
// closed form through dl = (a - d) / 2, b, c and t = (a + d) / 2, each affine in the pulse.
struct SatObj {
  cplx phi[2], eta[2];
  double d0, d1, b0, b1, c0, c1, t0, t1;
  double fixed;    // coefficient of term 1 when no pulse drives it
  bool driven;
};

struct SatPhi {
  cplx p0, p1;
};
// General step, out of line (by value: the objectives must stay in registers): generators
// with a trace (phase e^{it}), repeated scaled steps, any series degree.
static __device__ __noinline__ SatPhi sat_step_general(SatPhi v, double dl, double b, double c,
                                                       double t, int P, int s) {
  const double z = fma(dl, dl, b * c);
  double C = c_kq_tables.invfact[2 * (P - 1)], S = c_kq_tables.invfact[2 * (P - 1) + 1];
  for (int j = P - 2; j >= 0; --j) {
    C = fma(-z, C, c_kq_tables.invfact[2 * j]);
    S = fma(-z, S, c_kq_tables.invfact[2 * j + 1]);
  }
  const double Sd = S * dl, Sb = S * b, Sc = S * c;
  double st = 0.0, ct = 1.0;
  if (t != 0.0) sincos(t, &st, &ct);
  for (int rep = 0; rep < s; ++rep) {
    cplx u0 = make_double2(fma(-Sd, v.p0.y, fma(-Sb, v.p1.y, C * v.p0.x)),
                           fma(Sd, v.p0.x, fma(Sb, v.p1.x, C * v.p0.y)));
    cplx u1 = make_double2(fma(Sd, v.p1.y, fma(-Sc, v.p0.y, C * v.p1.x)),
                           fma(-Sd, v.p1.x, fma(Sc, v.p0.x, C * v.p1.y)));
    if (t != 0.0) {
      u0 = make_double2(fma(-st, u0.y, ct * u0.x), fma(st, u0.x, ct * u0.y));
      u1 = make_double2(fma(-st, u1.y, ct * u1.x), fma(st, u1.x, ct * u1.y));
    }
    v.p0 = u0;
    v.p1 = u1;
  }
  return v;
}

// phi <- exp(i R) phi for the thread's objectives TOGETHER, traceless generators and one
// unscaled step (the CTA-uniform common case): straight-line code with the series degree as a
// template parameter (coefficients are immediates, the independent chains of the objectives
// interleave).
template <int P>
__device__ __forceinline__ void sat_step_fast(SatObj (&o)[KQ_SAT_OPT], double h, double heps_new) {
  double dl[KQ_SAT_OPT], b[KQ_SAT_OPT], c[KQ_SAT_OPT], z[KQ_SAT_OPT];
  double C[KQ_SAT_OPT], S[KQ_SAT_OPT];
#pragma unroll
  for (int q = 0; q < KQ_SAT_OPT; ++q) {
    const double heps = o[q].driven ? heps_new : h * o[q].fixed;
    dl[q] = fma(heps, o[q].d1, h * o[q].d0);
    b[q] = fma(heps, o[q].b1, h * o[q].b0);
    c[q] = fma(heps, o[q].c1, h * o[q].c0);
    z[q] = fma(dl[q], dl[q], b[q] * c[q]);
    C[q] = kq_inv_fact(2 * (P - 1));
    S[q] = kq_inv_fact(2 * (P - 1) + 1);
  }
#pragma unroll
  for (int j = P - 2; j >= 0; --j) {
#pragma unroll
    for (int q = 0; q < KQ_SAT_OPT; ++q) {
      C[q] = fma(-z[q], C[q], kq_inv_fact(2 * j));
      S[q] = fma(-z[q], S[q], kq_inv_fact(2 * j + 1));
    }
  }
#pragma unroll
  for (int q = 0; q < KQ_SAT_OPT; ++q) {
    const double Sd = S[q] * dl[q], Sb = S[q] * b[q], Sc = S[q] * c[q], Cq = C[q];
    const cplx v0 = o[q].phi[0], v1 = o[q].phi[1];
    o[q].phi[0] = make_double2(fma(-Sd, v0.y, fma(-Sb, v1.y, Cq * v0.x)),
                               fma(Sd, v0.x, fma(Sb, v1.x, Cq * v0.y)));
    o[q].phi[1] = make_double2(fma(Sd, v1.y, fma(-Sc, v0.y, Cq * v1.x)),
                               fma(-Sd, v1.x, fma(Sc, v0.x, Cq * v1.y)));
  }
}

// CTA = warp 0 (exchange) + warp 1 (TMA producer) + KQ_SAT_BT consumer threads.
// shared: red [2][16] | gpart [2][8] (+ pad) | mbar [RING] | empty [RING] | sdt, sg, ssl, sbound [KQ_NTC] |
//         splan [KQ_NTC] | ring [RING][kpc][2] | M [kpc + 1][4] (||chi|| mu^T, real; last row zero)
__global__ void __launch_bounds__(KQ_SAT_THREADS, 1) k_fwupd_sat(const KqSweepArgs a, int kpc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (a.cond_epoch && *reinterpret_cast<volatile int*>(a.status + 1) != (int)a.cond_epoch) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NCW = KQ_SAT_BT / 32;   // consumer warps
  const int K = a.K, NT = a.NT, nblk = gridDim.x;
  const int k0 = blockIdx.x * kpc;
  const int kcta = max(0, min(kpc, K - k0));   // objectives of this CTA (0 for idle CTAs)
  double* red = reinterpret_cast<double*>(smem_raw);            // [2][16]
  double* gpart = red + 32;                                     // [2][8] gathered sums per polling warp
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red + 56);       // [RING] row has landed
  uint64_t* empty = mbar + KQ_SAT_RING;                         // [RING] row has been read
  double* sdt = reinterpret_cast<double*>(empty + KQ_SAT_RING);
  double* sg = sdt + KQ_NTC;
  double* ssl = sg + KQ_NTC;
  double* sbound = ssl + KQ_NTC;
  unsigned char* splan = reinterpret_cast<unsigned char*>(sbound + KQ_NTC);
  cplx* ring = reinterpret_cast<cplx*>(splan + KQ_NTC);
  const size_t stage = (size_t)kpc * 2;
  double* sM = reinterpret_cast<double*>(ring + (size_t)KQ_SAT_RING * stage);
  const uint32_t row_bytes = (uint32_t)kcta * 2 * sizeof(cplx);
  const bool xwarp = warp < 2;                 // exchange / producer warp: no objectives
  const int ctid = tid - 64;                   // consumer index
  bool failed = false;

  if (tid == 0) {
    for (int st = 0; st < KQ_SAT_RING; ++st) {
      mbar_init(&mbar[st], 1);
      mbar_init(&empty[st], (uint32_t)KQ_SAT_BT);   // every consumer thread releases what it read
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 16) gpart[tid] = 0.0;
  // idle CTAs never fill the ring: their (weight-zero) threads must still read finite numbers
  if (kcta == 0)
    for (size_t i = tid; i < (size_t)KQ_SAT_RING * stage; i += KQ_SAT_THREADS) ring[i] = c_zero();
  __syncthreads();
  // backward states chi(t_r), r = 0..NT: step n needs row n (its overlap) -- row NT only feeds
  // the (unused) overlap prepared behind the last step, so that the loop has no tail
  auto issue_row = [&](int r) {
    const int st = r % KQ_SAT_RING;
    mbar_expect_tx(&mbar[st], row_bytes);
    bulk_g2s(ring + st * stage, a.X + ((size_t)r * K + k0) * 2, row_bytes, &mbar[st]);
  };
  if (tid == 0 && kcta > 0)
    for (int r = 0; r < KQ_SAT_RING && r <= NT; ++r) issue_row(r);

  // ---- this thread's objectives: local indices ctid and ctid + BT (padding threads shadow
  // objective 0 of the CTA with weight zero)
  SatObj o[KQ_SAT_OPT];
  int kl[KQ_SAT_OPT];
  double O0 = 0.0, O1 = 0.0, Oc = 0.0;
#pragma unroll
  for (int q = 0; q < KQ_SAT_OPT; ++q) {
    const int l = ctid + q * KQ_SAT_BT;
    const bool valid = !xwarp && l < kcta;
    kl[q] = valid ? l : 0;
    const int k = min(k0 + kl[q], K - 1);
    double T0[4], T1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      T0[e] = -a.ops[((size_t)k * 2 + 0) * 4 + e].x;   // f = -i: exp(i R), R = -H dt
      T1[e] = -a.ops[((size_t)k * 2 + 1) * 4 + e].x;
    }
    o[q].d0 = 0.5 * (T0[0] - T0[3]);
    o[q].d1 = 0.5 * (T1[0] - T1[3]);
    o[q].t0 = 0.5 * (T0[0] + T0[3]);
    o[q].t1 = 0.5 * (T1[0] + T1[3]);
    o[q].c0 = T0[1];
    o[q].c1 = T1[1];
    o[q].b0 = T0[2];
    o[q].b1 = T1[2];
    const double cnorm = valid ? a.chi_norms[k] : 0.0;
    o[q].phi[0] = a.state0[(size_t)k * 2 + 0];
    o[q].phi[1] = a.state0[(size_t)k * 2 + 1];
    if (valid)
      for (int e = 0; e < 4; ++e) sM[(size_t)l * 4 + e] = cnorm * a.mu[(size_t)k * 4 + e].x;
    const int t2p = a.term2pulse[k * 2 + 1];
    o[q].driven = (t2p == 0);
    o[q].fixed = (t2p == -1) ? 1.0 : 0.0;
    if (valid) {
      O0 = fmax(O0, a.op_norm[k * 2 + 0]);
      O1 = fmax(O1, o[q].driven ? a.op_norm[k * 2 + 1] : 0.0);
      Oc = fmax(Oc, o[q].driven ? 0.0 : o[q].fixed * a.op_norm[k * 2 + 1]);
    }
  }
  // padding threads read the zero matrix behind the last objective: their eta is zero
  if (tid < 4) sM[(size_t)kpc * 4 + tid] = 0.0;
  O0 = block_max(O0, red);
  O1 = block_max(O1, red);
  Oc = block_max(Oc, red);
  double trace = 0.0;
#pragma unroll
  for (int q = 0; q < KQ_SAT_OPT; ++q) trace = fmax(trace, fabs(o[q].t0) + fabs(o[q].t1));
  const bool has_trace = block_max(trace, red) > 0.0;
  const double lam = a.lambda_a[0];
  // eta = mu^dag chi ||chi|| of the row in ring stage `st`; padding threads (q-th objective
  // beyond kcta) get zero through `w`
  int km[KQ_SAT_OPT];
#pragma unroll
  for (int q = 0; q < KQ_SAT_OPT; ++q) km[q] = (!xwarp && ctid + q * KQ_SAT_BT < kcta) ? kl[q] : kpc;
  auto make_eta_q = [&](int st, int q, cplx (&dst)[KQ_SAT_OPT][2]) {
    const cplx* row = ring + st * stage + (size_t)kl[q] * 2;
    const cplx x0 = row[0], x1 = row[1];
    const double2 m0 = *reinterpret_cast<const double2*>(sM + (size_t)km[q] * 4);
    const double2 m1 = *reinterpret_cast<const double2*>(sM + (size_t)km[q] * 4 + 2);
    // eta_c = sum_r M[c][r] chi_r (element (r, c) of mu at c * 2 + r)
    dst[q][0] = make_double2(fma(m0.x, x0.x, m0.y * x1.x), fma(m0.x, x0.y, m0.y * x1.y));
    dst[q][1] = make_double2(fma(m1.x, x0.x, m1.y * x1.x), fma(m1.x, x0.y, m1.y * x1.y));
  };
  auto make_eta = [&](int st, cplx (&dst)[KQ_SAT_OPT][2]) {
#pragma unroll
    for (int q = 0; q < KQ_SAT_OPT; ++q) make_eta_q(st, q, dst);
  };
  if (warp == 1) {
    // ---- producer warp: row r goes into stage r % RING as soon as every consumer warp has
    // released the row that was there (data-driven: no barrier shared with the other warps)
    if (lane == 0 && kcta > 0) {
      for (int r = KQ_SAT_RING; r <= NT; ++r) {
        const int st = r % KQ_SAT_RING, use = r / KQ_SAT_RING;
        mbar_wait(&empty[st], (uint32_t)((use - 1) & 1));
        issue_row(r);
      }
    }
    return;
  }
  if (kcta > 0) mbar_wait(&mbar[0], 0);
  {
    cplx e0[KQ_SAT_OPT][2];
    make_eta(0, e0);
#pragma unroll
    for (int q = 0; q < KQ_SAT_OPT; ++q) {
      o[q].eta[0] = e0[q][0];
      o[q].eta[1] = e0[q][1];
    }
    if (warp >= 2) sat_mbar_arrive(&empty[0]);
  }
  double ga = 0.0;
  // optional per-phase cycle counts, kq_set_option("picard_timing", 1): thread 0 (exchange
  // warp) and thread 64 (first consumer) of CTA 0
  const bool timing = a.pic_timing && (tid == 0 || (tid == 64 && blockIdx.x == 0));
  long long tacc[4] = {0, 0, 0, 0}, tprev = 0, npoll = 0;
#define KQ_SAT_TICK(i)                  \
  if (timing) {                         \
    const long long now_ = clock64();   \
    tacc[i] += now_ - tprev;            \
    tprev = now_;                       \
  }

  for (int base = 0; base < NT; base += KQ_NTC) {
    const int len = min(KQ_NTC, NT - base);
    sat_bar_a();
    for (int i = (warp == 0 ? tid : tid - 32); i < len; i += KQ_SAT_BT + 32) {
      const double dti = a.dt[base + i], gi = a.pulses[base + i];
      sdt[i] = dti;
      sg[i] = gi;
      ssl[i] = a.shape[base + i] / lam;   // S/lambda as in optimize.py:474
      int s, m;
      double bound;
      plan_bound(dti * (fma(fabs(gi), O1, O0) + Oc), s, m, bound);
      sbound[i] = bound;
      splan[i] = (s == 1) ? (unsigned char)m : (unsigned char)0;
    }
    sat_bar_a();
    if (timing) tprev = clock64();
    if (warp == 0) {
      // ---- exchange warp: the CTA's sum goes into the mailbox of every CTA (a mailbox's lines
      // are polled by their owner only -- a slot array read by all CTAs serialises 148 readers
      // on every line)
      for (int j = 0; j < len; ++j) {
        const int n = base + j, par = n & 1;
        sat_bar_a();   // barrier A: the consumer warps' partial sums are in red[par]
        KQ_SAT_TICK(0)
        double acc = (lane < NCW) ? red[par * 16 + lane] : 0.0;
        acc = warp_allreduce_sum(acc);
        const uint32_t tag = a.tag_base + (uint32_t)n + 1u;
        KqSlot* box = a.slots + (size_t)par * nblk * nblk;
        for (int r = lane; r < nblk; r += 32)
          sat_slot_store(box + (size_t)r * nblk + blockIdx.x, acc, tag);
        KQ_SAT_TICK(1)
      }
      continue;
    }
    // ---- consumers --------------------------------------------------------------------
    for (int j = 0; j < len; ++j) {
      const int n = base + j, par = n & 1;
      // Im <chi| mu |phi> ||chi|| of this thread's objectives
      double val = 0.0;
#pragma unroll
      for (int q = 0; q < KQ_SAT_OPT; ++q)
        val += c_im_conj_mul(o[q].eta[0], o[q].phi[0]) + c_im_conj_mul(o[q].eta[1], o[q].phi[1]);
      val = warp_allreduce_sum(val);
      if (lane == 0) red[par * 16 + warp - 2] = val;
      sat_bar_a();   // barrier A
      KQ_SAT_TICK(0)
      // the next backward states are in shared memory by now: eta of the next step is formed
      // while the exchange warp gathers the sum
      {
        const int r = n + 1, st = r % KQ_SAT_RING;
        if (kcta > 0) mbar_wait(&mbar[st], (uint32_t)((r / KQ_SAT_RING) & 1));
      }
      // gather: the first KQ_SAT_NG consumer warps poll 32 slots of this CTA's mailbox each,
      // ONE load per lane (loads of one thread complete one after the other: five slots per
      // lane in the exchange warp cost 1.6k cycles per poll).  The first poll is issued between
      // the two halves of eta: it samples the mailbox about when the peers' reductions land,
      // and its ~1000 cycles pass under the second half
      const bool helper = warp - 2 < KQ_SAT_NG;
      const uint32_t tag = a.tag_base + (uint32_t)n + 1u;
      const int gq = (warp - 2) * 32 + lane;
      const bool gact = helper && gq < nblk;
      const KqSlot* gp = a.slots + (size_t)par * nblk * nblk + (size_t)blockIdx.x * nblk + (gact ? gq : 0);
      uint32_t lo = 0, t0 = 0, hi = 0, t1 = 0;
      cplx eta_next[KQ_SAT_OPT][2];
      make_eta_q((n + 1) % KQ_SAT_RING, 0, eta_next);
      if (helper)
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1)
                     : "l"(gp)
                     : "memory");
      make_eta_q((n + 1) % KQ_SAT_RING, 1, eta_next);
      sat_mbar_arrive(&empty[(n + 1) % KQ_SAT_RING]);   // row n + 1 is consumed
      KQ_SAT_TICK(1)
      if (helper) {
        double v = 0.0;
        if (!failed) {
          int spin = 0;
          for (; spin < (1 << 22); ++spin) {
            const bool ok = !gact || (t0 == tag && t1 == tag);
            v = gact ? __hiloint2double((int)hi, (int)lo) : 0.0;
            if (__all_sync(0xffffffffu, ok)) break;
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1)
                         : "l"(gp)
                         : "memory");
          }
          npoll += spin + 1;
          if (spin == (1 << 22)) failed = true;
        }
        v = warp_allreduce_sum(failed ? 0.0 : v);
        if (lane == 0) gpart[par * 8 + warp - 2] = v;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(KQ_SAT_BT) : "memory");   // barrier B (consumers)
      KQ_SAT_TICK(2)
      double d1 = gpart[par * 8];
#pragma unroll
      for (int w2 = 1; w2 < KQ_SAT_NG; ++w2) d1 += gpart[par * 8 + w2];
      if (blockIdx.x == 0 && tid == 64) {   // pulse update (optimize.py:471-477)
        const double sl = ssl[j];
        a.opt_pulses[n] = __dadd_rn(sg[j], __dmul_rn(sl, d1));
        ga = __dadd_rn(ga, __dmul_rn(__dmul_rn(sl, __dmul_rn(d1, d1)), sdt[j]));
      }
      const double dt_cur = sdt[j];
      const double eps_new = __dadd_rn(sg[j], __dmul_rn(ssl[j], d1));
      // forward step under the updated pulse (series degree planned from the guess pulse,
      // verified against the updated one)
      const double x_new = dt_cur * (fma(fabs(eps_new), O1, O0) + Oc);
      int s = 1, m = splan[j];
      if (m == 0 || !(x_new <= sbound[j])) {
        double bound;
        plan_bound(x_new, s, m, bound);
      }
      const int P = max(2, (m + 4) >> 1);   // 2P >= m + 3; P > 12 (scaled norm close to 1) runs 18
      const double h = (s == 1) ? dt_cur : dt_cur / (double)s;
      const double heps = h * eps_new;
      if (s == 1 && !has_trace) {
        // a few degrees only (more terms than needed cost two DFMA each; every variant is
        // straight-line code in the loop body, and the loop should stay in the instruction cache)
        if (P <= 3) sat_step_fast<3>(o, h, heps);
        else if (P <= 5) sat_step_fast<5>(o, h, heps);
        else if (P <= 8) sat_step_fast<8>(o, h, heps);
        else if (P <= 12) sat_step_fast<12>(o, h, heps);
        else sat_step_fast<18>(o, h, heps);
      } else {
#pragma unroll
        for (int q = 0; q < KQ_SAT_OPT; ++q) {
          const double hq = o[q].driven ? heps : h * o[q].fixed;
          SatPhi v;
          v.p0 = o[q].phi[0];
          v.p1 = o[q].phi[1];
          v = sat_step_general(v, fma(hq, o[q].d1, h * o[q].d0), fma(hq, o[q].b1, h * o[q].b0),
                               fma(hq, o[q].c1, h * o[q].c0), fma(hq, o[q].t1, h * o[q].t0), P, s);
          o[q].phi[0] = v.p0;
          o[q].phi[1] = v.p1;
        }
      }
#pragma unroll
      for (int q = 0; q < KQ_SAT_OPT; ++q) {
        o[q].eta[0] = eta_next[q][0];
        o[q].eta[1] = eta_next[q][1];
      }
      KQ_SAT_TICK(3)
    }
  }
  if (timing && blockIdx.x == 0) {
    long long* out = reinterpret_cast<long long*>(a.status + 16) + (tid == 0 ? 0 : 5);
    for (int i = 0; i < 4; ++i) out[i] = tacc[i];
    out[4] = npoll;
  }
  if (timing && tid == 0) {
    // every CTA's exchange-warp counters behind the mailboxes: [cta][4 phases + polls + smid]
    long long* all = reinterpret_cast<long long*>(a.slots + (size_t)2 * nblk * nblk) + (size_t)blockIdx.x * 6;
    for (int i = 0; i < 4; ++i) all[i] = tacc[i];
    all[4] = npoll;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    all[5] = smid;
  }
  if (failed) atomicExch(a.status, (int)-4);
  if (xwarp) return;
  if (blockIdx.x == 0 && tid == 64) a.g_a[0] = ga;
  if (a.stateT) {
#pragma unroll
    for (int q = 0; q < KQ_SAT_OPT; ++q) {
      const int l = ctid + q * KQ_SAT_BT;
      if (l < kcta) {
        a.stateT[(size_t)(k0 + l) * 2 + 0] = o[q].phi[0];
        a.stateT[(size_t)(k0 + l) * 2 + 1] = o[q].phi[1];
      }
    }
  }
}
