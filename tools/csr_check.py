"""Check the CSR kernel family (64 < N <= 1024, csrc/kq_csr.cuh) against the numpy oracle on
the two-transmon Liouville problem of notebook 06 and time one iteration."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krotov_b200 as krotov
from oracle import krotov_oracle as orc

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 40
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wl = krotov.workloads.two_transmon_gate(n_qubit=nq, nt=nt, T=400.0 * nt / 2000)
low = wl.lowered()
print("N =", len(low['psi0'][0]), "K =", wl.K, "L =", len(low['pulses']), "nt =", nt, flush=True)
t0 = time.time()
res = krotov.optimize_pulses(
    wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
    propagator=krotov.propagators.expm, chi_constructor=krotov.functionals.chis_re,
    iter_stop=iters, store_all_pulses=True)
torch.cuda.synchronize()
print("gpu: %.3f s for %d iterations (incl. set-up)" % (time.time() - t0, iters),
      res.iter_seconds_device, flush=True)
if '--no-oracle' not in sys.argv:
    t0 = time.time()
    rec = orc.optimize(low['terms'], low['psi0'], low['targets'], low['pulses'], low['shapes'],
                       low['lambdas'], low['tlist'], orc.chis_re, is_super=True,
                       iter_stop=iters, weights=wl.weights)
    print("oracle: %.1f s" % (time.time() - t0), flush=True)
    for it in range(1, iters + 1):
        for l in range(len(low['pulses'])):
            a = np.asarray(res.all_pulses[it][l]); b = rec[it]['optimized_pulses'][l]
            print("it %d pulse %d: max |gpu - oracle| / max|oracle pulse 0| = %.3e (update %.3e)" % (
                it, l, np.max(np.abs(a - b)) / np.max(np.abs(rec[it]['optimized_pulses'][0])),
                np.max(np.abs(b - low['pulses'][l]))))
