"""Iterations per second of the reference's notebook-shaped problems through the public API
(no hooks: all iterations queued, one synchronisation), with the CPU oracle beside them on a
bounded sample: Lambda system (notebooks 02/03: N = 3, four controls), its 5-member ensemble
(notebook 08), the transmon X gate with 5 and 17 levels (notebook 05)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krotov_b200 as krotov
from oracle import krotov_oracle as orc
from threadpoolctl import threadpool_limits
W = krotov.workloads
CASES = [
    ('lambda_rwa_nonherm (nb 03: N=3, L=4, K=1, nt=500)', lambda: W.lambda_system(nt=500, gamma=0.5)),
    ('lambda_ensemble (nb 08: N=3, L=4, K=5, nt=500)',
     lambda: W.lambda_system(nt=500, gamma=0.0, lambda_a=0.5, ensemble_mu=[0.9, 0.95, 1.0, 1.05, 1.1])),
    ('transmon_xgate N=5 (K=2, nt=1000)', lambda: W.transmon_xgate(nstates=2, nt=1000)),
    ('transmon_xgate N=17 (nb 05: K=2, nt=1000)', lambda: W.transmon_xgate(nstates=8, nt=1000)),
]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
only = sys.argv[2] if len(sys.argv) > 2 else ''   # substring filter on the case name
for name, make in CASES:
    if only not in name:
        continue
    wl = make()
    chi = getattr(krotov.functionals, 'chis_' + wl.chi)
    def run(n, **kw):
        return krotov.optimize_pulses(
            wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
            propagator=krotov.propagators.expm, chi_constructor=chi, iter_stop=n, **kw)
    run(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); r1 = run(5); torch.cuda.synchronize(); t1 = time.perf_counter()
    r2 = run(5 + iters); torch.cuda.synchronize(); t2 = time.perf_counter()
    gpu = iters / ((t2 - t1) - (t1 - t0))
    # the same with a host hook per iteration (as the reference's notebooks run: print_table)
    stamps = []
    run(5 + iters, info_hook=lambda **kw: stamps.append(time.perf_counter()))
    hooked = iters / (stamps[-1] - stamps[-1 - iters])
    low = wl.lowered()
    chi_o = getattr(orc, 'chis_' + wl.chi)
    with threadpool_limits(1):
        t0 = time.perf_counter()
        orc.optimize(low['terms'], low['psi0'], low['targets'], low['pulses'], low['shapes'],
                     low['lambdas'], low['tlist'], chi_o, is_super=low['is_super'], iter_stop=1,
                     weights=low['weights'])
        cpu = time.perf_counter() - t0
    # iteration 0 (forward propagation) is about a third of the oracle's 1-iteration run
    print("%-52s GPU %8.1f it/s (with a hook per iteration %8.1f)   CPU port (1 core) %6.2f it/s   fused=%s launches/it=%.1f" % (
        name, gpu, hooked, 1.0 / (cpu * 2.0 / 3.0), getattr(r2, 'fused_iterations', None),
        (r2.gpu_launches - r1.gpu_launches) / iters), flush=True)
