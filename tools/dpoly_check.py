"""Check and time the delta-polynomial Krotov iteration (kq_krotov_iteration on the
composite path, csrc/kq_dpoly.cuh) against the sequential Taylor sweep kernels through
the C ABI: pulses, phi(T), tau, backward states per iteration, and the plan header
(degree, builds / reuses)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krotov_b200 as krotov
from krotov_b200.compiler import compile_problem, initialize_controls
from krotov_b200.engine import SweepEngine

lib = krotov._lib.load()
W = krotov.workloads
WL = {'C1': lambda: W.tls_state_to_state(),
      'C2': lambda: W.transmon_xgate(),
      'C3': lambda: W.two_qubit_gate(nt=2000),
      'C5': lambda: W.dissipative_qubit_reset(nt=5000),
      'C5s': lambda: W.dissipative_qubit_reset(nt=500),
      'T5': lambda: W.transmon_xgate(nstates=2, nt=1000),
      'K8': lambda: W.tls_ensemble(K=8, nt=1000)}


def header(eng):
    off = lib.kq_dpoly_header_offset(eng._p)
    h = eng.workspace[off:off + 64].cpu().numpy()
    i = h[:32].view(np.int32)
    d = h[32:].view(np.float64)
    return dict(J=int(i[0]), m=int(i[1]), rebuild=int(i[2]), usable=int(i[3]),
                builds=int(i[6]), reuses=int(i[7]), radius=float(d[0]),
                last_max=float(d[1]))


def run(name, iters):
    wl = WL[name]()
    objectives = wl.objectives(krotov.Objective)
    controls, _, guess, mapping, lam, shp = initialize_controls(
        objectives, wl.pulse_options, wl.tlist)
    cp = compile_problem(objectives, controls, mapping, wl.tlist)
    chi_kind = None if wl.chi == 'qubit_reset' else wl.chi

    def sweeps():
        lib.kq_set_option(b"dpoly", 0)
        lib.kq_set_option(b"picard", 0)
        eng = SweepEngine(cp, shp, lam)
        g = eng.pulses_to_device(guess)
        o = g.clone()
        phiT = eng.propagate_forward(g)
        tau = eng.overlaps(eng.t_targets, phiT)
        out = []
        for it in range(iters):
            if chi_kind is None:
                eng.chi_from_host([cp.vec(wl.meta['chi_fixed'])] * cp.K)
            else:
                eng.chi_builtin(chi_kind, phiT, tau)
            eng.sweep_backward(g)
            phiT = eng.sweep_forward_update(g, o, phiT=eng.new_states())
            tau = eng.overlaps(eng.t_targets, phiT)
            torch.cuda.synchronize()
            out.append((o.cpu().numpy().copy(), phiT.cpu().numpy().copy(),
                        tau.cpu().numpy().copy(), eng.X.cpu().numpy().copy(),
                        eng.g_a.cpu().numpy().copy()))
            g, o = o, g
        return out

    def composite():
        lib.kq_set_option(b"dpoly", 2)
        lib.kq_set_option(b"picard", 1)
        eng = SweepEngine(cp, shp, lam)
        g = eng.pulses_to_device(guess)
        o = g.clone()
        phiT = eng.propagate_forward(g)
        tau = eng.overlaps(eng.t_targets, phiT)
        out = []
        diag = torch.zeros(4, dtype=torch.int32, device=eng.device)
        for it in range(iters):
            if chi_kind is None:
                eng.chi_from_host([cp.vec(wl.meta['chi_fixed'])] * cp.K)
            p2, t2 = eng.new_states(), torch.empty_like(tau)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.krotov_iteration(chi_kind, g, o, phiT, tau, p2, t2, store_X=True,
                                 diag_t=diag)
            e1.record()
            torch.cuda.synchronize()
            phiT, tau = p2, t2
            out.append((o.cpu().numpy().copy(), phiT.cpu().numpy().copy(),
                        tau.cpu().numpy().copy(), eng.X.cpu().numpy().copy(),
                        eng.g_a.cpu().numpy().copy(), header(eng),
                        diag.cpu().numpy().copy(), e0.elapsed_time(e1)))
            g, o = o, g
        return out

    ref = sweeps()
    got = composite()
    worst = 0.0
    for it, (a, b) in enumerate(zip(ref, got)):
        sc = np.max(np.abs(a[0]))
        err = np.max(np.abs(a[0] - b[0])) / sc
        worst = max(worst, err)
        print("%s it %d: pulse %.2e phiT %.2e tau %.2e X %.2e g_a %.2e | %.1f us | "
              "diag %s hdr %s" % (
                  name, it + 1, err, np.max(np.abs(a[1] - b[1])),
                  np.max(np.abs(a[2] - b[2])), np.max(np.abs(a[3] - b[3])),
                  np.max(np.abs(a[4] - b[4])) / max(np.max(np.abs(a[4])), 1e-300),
                  b[7] * 1e3, b[6].tolist(), b[5]), flush=True)
    return worst


if __name__ == '__main__':
    names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['C2', 'C3', 'C5s']
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    bad = 0
    for nm in names:
        w = run(nm, iters)
        print("%s worst rel pulse err %.3e" % (nm, w))
        bad += w > 1e-10
    sys.exit(1 if bad else 0)
