"""TEST / BENCH INFRASTRUCTURE ONLY -- run the UNMODIFIED reference package
(``krotov``) on a workload and time its iterations.

The reference is imported from ``/root/reference/src`` (this container) or
from the git-ignored copy ``baseline/_ref`` that ``__graft_entry__.build()``
places there (GPU box), through the stand-ins in ``oracle/ref_shims`` (dense
numpy-backed ``qutip.Qobj``, ``glom``, ``grapheme``; QuTiP itself cannot be
installed here) plus the NumPy-2 alias ``np.ComplexWarning``.  The
reference's own ``optimize_pulses`` loop, ``propagators.expm``,
``mu.derivative_wrt_pulse`` and ``functionals.chis_*`` then run as written
(optimize.py:33-590).

Only ``bench.py`` (cpu_baseline legs / ``--impl reference``) and the tests use
this module.  It runs in a process of its own (``python -m
oracle.run_reference ...``) so that the name ``krotov`` and the shims never
enter the process that holds the CUDA engine.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_path():
    for cand in ('/root/reference/src', os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(cand, 'krotov', 'optimize.py')):
            return cand
    return None


def import_reference():
    path = reference_path()
    if path is None:
        raise ImportError("the reference package is neither at "
                          "/root/reference/src nor in baseline/_ref")
    for p in (path, os.path.join(HERE, 'ref_shims'), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, 'ComplexWarning'):
        np.ComplexWarning = np.exceptions.ComplexWarning
    import krotov
    import qutip
    return krotov, qutip


def _numpy_plugins(is_super):
    """User plugins of the reference's own numpy recipe
    (docs/notebooks/09_example_numpy.ipynb cells 16/30/32): propagator, mu and
    overlap on plain arrays; density matrices are column-stacked."""
    import scipy.linalg

    def vec(s):
        return s.reshape(-1, order='F') if is_super else s

    def unvec(v, like):
        return v.reshape(like.shape, order='F') if is_super else v

    def expm(H, state, dt, c_ops=None, backwards=False, initialize=False):
        f = 1.0 if is_super else (1j if backwards else -1j)
        A = f * H[0]
        for part in H[1:]:
            A = A + (f * part[1]) * part[0]
        return unvec(scipy.linalg.expm(A * dt) @ vec(state), state)

    def overlap(a, b):
        if a is None or b is None:
            return None
        return complex(np.vdot(vec(a), vec(b)))

    def mu(objectives, i_objective, pulses, pulses_mapping, i_pulse,
           time_index):
        idx = pulses_mapping[i_objective][0][i_pulse]
        f = 1j if is_super else 1.0

        def _mu(state):
            out = 0 * vec(state)
            for i in idx:
                out = out + f * (objectives[i_objective].H[i][0] @ vec(state))
            return unvec(out, state)
        return _mu

    return dict(propagator=expm, overlap=overlap, mu=mu,
                norm=np.linalg.norm)


def time_reference(workload, k_sample, iters, kwargs=None, path='numpy'):
    """Seconds per Krotov iteration of the reference's serial loop
    (``parallel_map=None``; BLAS pinned to one thread by optimize.py:233-238)
    on the first `k_sample` objectives of the named workload.  `path`:
    'numpy' = plain arrays with the plugins of notebook 09 (the faster way to
    run the reference), 'qobj' = Qobj stand-ins and ``krotov.propagators.expm``."""
    krotov, qutip = import_reference()
    from krotov_b200 import workloads
    wl = getattr(workloads, workload)(**(kwargs or {}))
    K = wl.K
    ks = max(1, min(K, k_sample))
    krotov.Objective.type_checking = (path == 'qobj')

    def wrap(a):
        if path != 'qobj':
            return a
        if wl.is_super and a.shape == (16, 16):
            return qutip.Qobj(a, dims=[[[4], [4]], [[4], [4]]])
        return qutip.Qobj(a)

    objectives = wl.objectives(krotov.Objective, wrap=wrap)[:ks]
    if wl.chi == 'qubit_reset':
        fixed = wl.meta['chi_fixed']

        def chi_constructor(fw_states_T, objectives, tau_vals):
            return [wrap(fixed.copy()) for _ in fw_states_T]
    elif path == 'qobj':
        chi_constructor = getattr(krotov.functionals, 'chis_' + wl.chi)
    else:
        from oracle import krotov_oracle as orc
        fn = getattr(orc, 'chis_' + wl.chi)

        def chi_constructor(fw_states_T, objectives, tau_vals):
            out = fn([s.reshape(-1) for s in fw_states_T],
                     [o.target.reshape(-1) for o in objectives],
                     list(tau_vals), None)
            # the prefactors 1/N of functionals.py use the number of
            # objectives handed to the call
            return [c.reshape(o.target.shape)
                    for c, o in zip(out, objectives)]
    plug = dict(propagator=krotov.propagators.expm) if path == 'qobj' \
        else _numpy_plugins(wl.is_super)
    stamps, pulses = [], []

    def hook(**kw):
        stamps.append((kw['start_time'], kw['stop_time']))
        pulses.append(np.array([p.copy() for p in kw['optimized_pulses']]))
        return None

    t0 = time.perf_counter()
    krotov.optimize_pulses(
        objectives, pulse_options=wl.pulse_options, tlist=wl.tlist,
        chi_constructor=chi_constructor, info_hook=hook, iter_stop=iters,
        **plug)
    total = time.perf_counter() - t0
    per_iter = float(np.mean([b - a for a, b in stamps[1:]]))
    return dict(seconds_per_iteration_sample=per_iter, k_sample=ks, K=K,
                seconds_per_iteration=per_iter * K / ks, iterations=iters,
                total_seconds=total, reference_path=reference_path(),
                path=path, pulses_last=pulses[-1].tolist(),
                nt=len(wl.tlist))


if __name__ == '__main__':
    # python -m oracle.run_reference <workload> <k_sample> <iters> [json kwargs]
    name, ks, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    kw = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
    print(json.dumps(time_reference(name, ks, iters, kw)))
