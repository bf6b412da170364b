"""How much of the per-iteration time is host launch overhead?  Runs the fused
iteration back to back without synchronisation and reports GPU time per
iteration (events) and host time per call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krotov_b200 as krotov
from krotov_b200.compiler import compile_problem, initialize_controls
from krotov_b200.engine import SweepEngine

if os.environ.get('KQ_COOP') is not None:
    krotov._lib.load().kq_set_option(b"cooperative_launch", int(os.environ['KQ_COOP']))
if os.environ.get('KQ_PDL') is not None:
    krotov._lib.load().kq_set_option(b"programmatic_launch", int(os.environ['KQ_PDL']))
wl = krotov.workloads.tls_ensemble(K=128, nt=1000)
objectives = wl.objectives(krotov.Objective)
(controls, _, guess_pulses, mapping, lam, shp) = initialize_controls(objectives, wl.pulse_options, wl.tlist)
cp = compile_problem(objectives, controls, mapping, wl.tlist)
eng = SweepEngine(cp, shp, lam)
guess_t = eng.pulses_to_device(guess_pulses)
opt_t = guess_t.clone()
phiT = eng.propagate_forward(guess_t)
tau_t = eng.overlaps(eng.t_targets, phiT)
phiT2, tau2 = eng.new_states(), torch.empty_like(tau_t)
stream = torch.cuda.current_stream()

def run(n, hint=True):
    global guess_t, opt_t, phiT, phiT2, tau_t, tau2
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(n):
        eng.krotov_iteration('re', guess_t, opt_t, phiT, tau_t, phiT2, tau2, prev_guess_t=opt_t if hint else None)
        phiT, phiT2 = phiT2, phiT
        tau_t, tau2 = tau2, tau_t
        guess_t, opt_t = opt_t, guess_t
    e1.record(stream)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n, (t1 - t0) * 1e6 / n

for n in (1, 5, 50, 200):
    g, h = run(n)
    print("n=%3d  GPU us/iteration %.1f   host us/call %.1f   picard rounds %d" % (n, g, h, eng.sweep_diagnostics()[1]))
# restart from the guess for comparable convergence stage
