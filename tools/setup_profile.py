"""Where does the fixed cost of one optimize_pulses call go?"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krotov_b200 as krotov
torch.cuda.init(); torch.zeros(1, device='cuda')
wl = krotov.workloads.tls_ensemble(K=128, nt=1000)
def run():
    return krotov.optimize_pulses(
        wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
        propagator=krotov.propagators.expm, chi_constructor=krotov.functionals.chis_re,
        info_hook=lambda **kw: None, iter_stop=3)
t0 = time.perf_counter(); run(); t1 = time.perf_counter(); run(); t2 = time.perf_counter()
print("first call %.1f ms, second call %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
