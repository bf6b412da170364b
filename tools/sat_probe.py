"""Where a time step of the many-objective update sweep (csrc/kq_sat.cuh) goes: per-phase cycle
counts of thread 0 of CTA 0 (kq_set_option("picard_timing", 1)) and event-timed sweeps for
several ensemble sizes.  Usage: python tools/sat_probe.py [K ...]"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krotov_b200 as krotov
from krotov_b200.compiler import compile_problem, initialize_controls
from krotov_b200.engine import SweepEngine

lib = krotov._lib.load()
Ks = [int(v) for v in sys.argv[1:] if '=' not in v] or [9472, 32768, 131072]

for v in sys.argv[1:]:
    if v.startswith('kpc='):   # fewer, fuller CTAs: exchange latency against the number of participants
        lib.kq_set_option(b"sat_min_kpc", int(v[4:]))
        KPC = int(v[4:])
XNAMES = ['wait for barrier A', 'CTA reduce + push', '-', '-']
CNAMES = ['overlap + warp reduce + barrier A', 'next eta', 'gather + barrier B', 'update + step']
for K in Ks:
    wl = krotov.workloads.tls_ensemble(K=K, nt=1000)
    objectives = wl.objectives(krotov.Objective)
    (controls, _, guess, mapping, lam, shp) = initialize_controls(objectives, wl.pulse_options, wl.tlist)
    cp = compile_problem(objectives, controls, mapping, wl.tlist)
    eng = SweepEngine(cp, shp, lam)
    g = eng.pulses_to_device(guess)
    o = g.clone()
    phiT = eng.propagate_forward(g)
    tau = eng.overlaps(eng.t_targets, phiT)
    lib.kq_set_option(b"picard_timing", 1)
    st = torch.cuda.current_stream()
    ms_fw, ms_bw = [], []
    for it in range(4):
        eng.chi_builtin(wl.chi, phiT, tau, K_total=K)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(st)
        eng.sweep_backward(g)
        e[1].record(st)
        phiT = eng.sweep_forward_update(g, o, phiT=phiT)
        e[2].record(st)
        e[2].synchronize()
        ms_bw.append(e[0].elapsed_time(e[1]))
        ms_fw.append(e[1].elapsed_time(e[2]))
        tau = eng.overlaps(eng.t_targets, phiT)
        g, o = o, g
    lib.kq_set_option(b"picard_timing", 0)
    cyc = eng.workspace[64:64 + 80].view(torch.int64).cpu().numpy()
    polls, xc, cc = cyc[9], cyc[:4], cyc[5:9]
    NT = cp.NT
    print("K=%d: backward %.3f ms, update sweep %.3f ms = %.0f ns / step; cycles per step, exchange warp: %s (sum %d, "
          "polls per step %.2f); first consumer: %s (sum %d)" % (
        K, min(ms_bw), min(ms_fw), min(ms_fw) * 1e6 / NT,
        ', '.join('%s %d' % (n, c // NT) for n, c in zip(XNAMES, xc)), sum(xc) // NT, polls / NT,
        ', '.join('%s %d' % (n, c // NT) for n, c in zip(CNAMES, cc)), sum(cc) // NT), flush=True)
    # every CTA's exchange warp (behind the mailboxes in the workspace's slot area)
    kpc_ = max(globals().get('KPC', 64), -(-K // 148))
    kpc_ += kpc_ & 1
    nblk = (K + kpc_ - 1) // kpc_
    off = 256 + 2 * nblk * nblk * 16
    allc = eng.workspace[off:off + nblk * 48].view(torch.int64).cpu().numpy().reshape(nblk, 6)
    per = allc[:, :4] / NT
    order = np.argsort(per[:, 1])
    print("  per CTA (cycles/step): wait-A+reduce+store min %.0f median %.0f max %.0f | gather min %.0f median %.0f max %.0f | polls/step min %.2f max %.2f" % (
        per[:, 1].min(), np.median(per[:, 1]), per[:, 1].max(), per[:, 2].min(), np.median(per[:, 2]), per[:, 2].max(),
        allc[:, 4].min() / NT, allc[:, 4].max() / NT))
    print("  CTAs with the SHORTEST wait at barrier A (= the slowest consumers): " + ", ".join(
        "cta %d sm %d: %.0f/%.0f" % (i, allc[i, 5], per[i, 1], per[i, 2]) for i in order[:6]))
    print("  CTAs with the LONGEST wait at barrier A: " + ", ".join(
        "cta %d sm %d: %.0f/%.0f" % (i, allc[i, 5], per[i, 1], per[i, 2]) for i in order[-6:]), flush=True)
    del eng
    torch.cuda.empty_cache()
