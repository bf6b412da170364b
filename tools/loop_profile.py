"""Where does the host time of one hooked Krotov iteration go?  cProfile of
optimize_pulses (C4, info_hook, 400 iterations), sorted by own time."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krotov_b200 as krotov
torch.cuda.init(); torch.zeros(1, device='cuda')
wl = krotov.workloads.tls_ensemble(K=128, nt=1000)
N_IT = int(sys.argv[1]) if len(sys.argv) > 1 else 400
stamps = []
def hook(**kw):
    stamps.append(time.perf_counter())
    return 1 - np.mean(kw['tau_vals']).real
def run():
    stamps.clear()
    return krotov.optimize_pulses(
        wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
        propagator=krotov.propagators.expm, chi_constructor=krotov.functionals.chis_re,
        info_hook=hook, iter_stop=N_IT)
run()
run()
d = np.diff(stamps)[N_IT // 2:]
print("unprofiled: %.1f us per iteration (median %.1f)" % (d.mean() * 1e6, np.median(d) * 1e6))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(28)
