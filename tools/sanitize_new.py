"""Small runs of the round-2 kernel families for compute-sanitizer (memcheck / racecheck):
entries-in-registers kernels (csrc/kq_lanes.cuh: Lambda system with four controls, transmon
N = 17 in two warps) and the many-objective update sweep (csrc/kq_sat.cuh, K = 2300)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krotov_b200 as krotov

W = krotov.workloads
for wl in (W.lambda_system(nt=60, gamma=0.5),
           W.lambda_system(nt=60, gamma=0.0, lambda_a=0.5, ensemble_mu=[0.9, 0.95, 1.0, 1.05, 1.1]),
           W.transmon_xgate(nstates=8, nt=70),
           W.tls_ensemble(K=2300, nt=24)):
    chi = getattr(krotov.functionals, 'chis_' + wl.chi)
    res = krotov.optimize_pulses(
        wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
        propagator=krotov.propagators.expm, chi_constructor=chi,
        info_hook=lambda **kw: None, iter_stop=2)
    print(wl.name, 'fused iterations', res.fused_iterations, 'launches', res.gpu_launches,
          'max pulse', float(np.max(np.abs(res.optimized_controls[0]))), flush=True)
