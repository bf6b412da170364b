"""Windowed update sweep: time and total fixed-point rounds vs number of windows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krotov_b200 as krotov
from krotov_b200.compiler import compile_problem, initialize_controls
from krotov_b200.engine import SweepEngine

name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
W = krotov.workloads
wl = {'C2': W.transmon_xgate, 'C3': W.two_qubit_gate, 'C4': lambda: W.tls_ensemble(K=128, nt=1000),
      'C1': W.tls_state_to_state}[name]()
objectives = wl.objectives(krotov.Objective)
(controls, _, guess_pulses, mapping, lam, shp) = initialize_controls(objectives, wl.pulse_options, wl.tlist)
cp = compile_problem(objectives, controls, mapping, wl.tlist)
stream = torch.cuda.current_stream()
ref = None
for nwin in (1, 2, 3, 4, 6, 8, 16):
    eng = SweepEngine(cp, shp, lam)
    eng.problem.reserved = nwin
    guess_t = eng.pulses_to_device(guess_pulses); opt_t = guess_t.clone()
    phiT = eng.propagate_forward(guess_t); tau_t = eng.overlaps(eng.t_targets, phiT)
    ms, rounds = [], []
    for it in range(4):
        eng.chi_builtin(wl.chi if wl.chi in ('re', 'ss', 'sm', 'hs') else 're', phiT, tau_t, K_total=cp.K)
        eng.sweep_backward(guess_t)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        phiT = eng.sweep_forward_update(guess_t, opt_t, phiT=phiT)
        e1.record(stream)
        tau_t = eng.overlaps(eng.t_targets, phiT)
        torch.cuda.synchronize()
        fb, r = eng.sweep_diagnostics()
        ms.append(e0.elapsed_time(e1)); rounds.append(r)
        guess_t, opt_t = opt_t, guess_t
    pulses = guess_t.cpu().numpy()
    if ref is None:
        ref = pulses
    print("%s windows %2d: fw ms %s  total rounds %s  dev vs 1 window %.1e  fallback %s" % (
        name, nwin, ' '.join('%.3f' % m for m in ms), rounds, np.max(np.abs(pulses - ref)) / np.max(np.abs(ref)), fb == eng.epoch))
