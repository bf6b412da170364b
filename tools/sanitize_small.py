"""Small fused-iteration runs for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krotov_b200 as krotov

for wl, chi in ((krotov.workloads.tls_ensemble(K=6, nt=70), krotov.functionals.chis_sm),
                (krotov.workloads.transmon_xgate(nt=60), krotov.functionals.chis_re),
                (krotov.workloads.tls_state_to_state(nt=90), krotov.functionals.chis_ss)):
    res = krotov.optimize_pulses(
        wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
        propagator=krotov.propagators.expm, chi_constructor=chi,
        info_hook=lambda **kw: None, iter_stop=2)
    print(wl.name, 'fused iterations', res.fused_iterations, 'max pulse', float(np.max(np.abs(res.optimized_controls[0]))))
