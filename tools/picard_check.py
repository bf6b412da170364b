"""GPU check of the time-parallel (Picard) fused sweep against the sequential
kernel: same pulses to rounding, iteration counts, kernel times."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krotov_b200 as krotov
from krotov_b200.compiler import compile_problem, initialize_controls
from krotov_b200.engine import SweepEngine

lib = krotov._lib.load()


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def engine_run(wl, picard, iters=3, second=False, fused=False):
    lib.kq_set_option(b"picard", picard)
    objectives = wl.objectives(krotov.Objective)
    (controls, _, guess_pulses, mapping, lam, shp) = initialize_controls(
        objectives, wl.pulse_options, wl.tlist)
    cp = compile_problem(objectives, controls, mapping, wl.tlist)
    eng = SweepEngine(cp, shp, lam)
    guess_t = eng.pulses_to_device(guess_pulses)
    opt_t = guess_t.clone()
    Phi0 = eng.new_state_store() if second else None
    Phi1 = eng.new_state_store() if second else None
    phiT = eng.propagate_forward(guess_t, store=Phi0)
    tau_t = eng.overlaps(eng.t_targets, phiT)
    sigma_t = None
    if second:
        sigma_t = torch.full((cp.NT,), -0.05, dtype=torch.float64, device=eng.device)
    out = []
    stream = torch.cuda.current_stream()
    chi = wl.chi if wl.chi in ('re', 'ss', 'sm', 'hs') else 're'
    phiT2, tau2 = eng.new_states(), torch.empty_like(tau_t)
    for it in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        if fused:
            e0.record(stream)
            eng.krotov_iteration(chi, guess_t, opt_t, phiT, tau_t, phiT2, tau2, store_X=True,
                                 sigma_t=sigma_t, Phi0=Phi0, Phi1=Phi1, prev_guess_t=opt_t if it > 0 else None)
            e1.record(stream)
            phiT, phiT2 = phiT2, phiT
            tau_t, tau2 = tau2, tau_t
        else:
            eng.chi_builtin(chi, phiT, tau_t, K_total=cp.K)
            eng.sweep_backward(guess_t)
            e0.record(stream)
            phiT = eng.sweep_forward_update(guess_t, opt_t, phiT=phiT, sigma_t=sigma_t, Phi0=Phi0, Phi1=Phi1)
            e1.record(stream)
            tau_t = eng.overlaps(eng.t_targets, phiT)
        torch.cuda.synchronize()
        fb, pit = eng.sweep_diagnostics()
        cyc = eng.workspace[64:128].view(torch.int32).cpu().numpy().copy()
        out.append(dict(pulses=opt_t.cpu().numpy().copy(), phiT=phiT.cpu().numpy().copy(),
                        ga=eng.g_a.cpu().numpy().copy(), tau=tau_t.cpu().numpy().copy(),
                        ms=e0.elapsed_time(e1), cycles=cyc, fallback=(fb == eng.epoch), pit=pit,
                        Phi1=None if Phi1 is None else Phi1.cpu().numpy().copy(),
                        X=eng.X.cpu().numpy().copy(), chi=eng.chi.cpu().numpy().copy(),
                        status=eng.status()))
        guess_t, opt_t = opt_t, guess_t
        if second:
            Phi0, Phi1 = Phi1, Phi0
    return out


def compare(name, wl, second=False, iters=3):
    seq = engine_run(wl, 0, iters, second)
    pic = engine_run(wl, 2, iters, second)
    try:
        fus = engine_run(wl, 2, iters, second, fused=True)
    except Exception as exc:
        print(name, 'fused path unavailable:', exc)
        fus = None
    for label, run in (('picard', pic), ('fused ', fus)):
        if run is None:
            continue
        for it in range(iters):
            a, b = run[it], seq[it]
            line = ("%-18s %s it%d  pulses %.2e  phiT %.2e  ga %.2e  tau %.2e X %.2e chi %.2e | its %3d fb %d "
                    "st %d | ms: %.4f  seq-fw %.4f" % (
                        name, label, it + 1, rel(a['pulses'], b['pulses']),
                        np.max(np.abs(a['phiT'] - b['phiT'])), rel(a['ga'], b['ga']),
                        np.max(np.abs(a['tau'] - b['tau'])), np.max(np.abs(a['X'] - b['X'])),
                        np.max(np.abs(a['chi'] - b['chi'])), a['pit'], a['fallback'], a['status'],
                        a['ms'], b['ms']))
            if second:
                line += "  Phi1 %.2e" % np.max(np.abs(a['Phi1'] - b['Phi1']))
            print(line, flush=True)


def timing(name, wl, iters=3, fused=True):
    lib.kq_set_option(b"picard_timing", 1)
    out = engine_run(wl, 1, iters, fused=fused)
    lib.kq_set_option(b"picard_timing", 0)
    return out


if __name__ == '__main__':
    W = krotov.workloads
    if len(sys.argv) > 1 and sys.argv[1] == 'timing':
        name = sys.argv[2] if len(sys.argv) > 2 else 'C4'
        wl = {'C4': lambda: W.tls_ensemble(K=128, nt=1000), 'C1': W.tls_state_to_state,
              'C2': W.transmon_xgate, 'C3': W.two_qubit_gate}[name]()
        for r in timing(name, wl, 40)[-3:]:
            cyc = r['cycles']
            its = max(r['pit'], 1)
            names = ['prologue', 'bw', 'passA', 'passB', 'stage1', 'stage2', 'stage3', 'entry-to-exit-ns', 'outputs', 'final',
                     'scan-shfl', 'scan-bar', 'scan-finish', 'passB-bar', 'fetch', 'entry-to-exit']
            print('kernel %.1f us, %d its; cycles (per iter except prologue/bw/outputs/final): ' % (r['ms'] * 1e3, its) + ', '.join(
                '%s %d' % (n, c // (1 if i in (0, 1, 7, 8, 9, 15) else its)) for i, (n, c) in enumerate(zip(names, cyc)) if n != '-') + '  => SM clock %.3f GHz' % (cyc[15] / max(cyc[7], 1)))
        sys.exit(0)
    compare('C4 K=8 nt=100', W.tls_ensemble(K=8, nt=100))
    compare('C4 K=128 nt=1000', W.tls_ensemble(K=128, nt=1000), iters=5)
    compare('C1', W.tls_state_to_state())
    compare('C2', W.transmon_xgate())
    compare('C3 first', W.two_qubit_gate())
    compare('C3 second', W.two_qubit_gate(), second=True)
    compare('C4 K=300 nt=1000', W.tls_ensemble(K=300, nt=1000))
    compare('C4 K=128 nt=5000', W.tls_ensemble(K=128, nt=5000))
    lib.kq_set_option(b"picard", 1)
