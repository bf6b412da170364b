"""TEST INFRASTRUCTURE ONLY -- multi-process CPU oracle (all host cores).

Restatement of the reference's *parallel* mode of the hot path,

    parallel_map=(parallel_map, parallel_map, parallel_map_fw_prop_step)

(/root/reference/src/krotov/parallelization.py:51-57): the backward sweep
runs in parallel over the objectives without communication (:16-35), the
update/forward sweep with long-running consumer processes that synchronise
once **per time step** (:36-49, `Consumer` :314-354, `FwPropStepTask`
:357-430, `parallel_map_fw_prop_step` :433-495).  The arithmetic is the
serial oracle's (oracle/krotov_oracle.py: same `expm_step`,
`backward_propagation`, `mu_operator`, chi constructors), so the pulses
agree with the serial oracle to summation order (the partial sums over the
objectives are formed per worker, then added in worker order).

Differences from the reference's process layout, all in the CPU's favour:
one worker per *block* of objectives (the reference starts one consumer per
objective, :472-479), the workers keep their backward states and send one
complex partial sum per pulse and step instead of pickled states, and pipes
replace the JoinableQueues.

Only ``tests/`` and ``bench.py``'s CPU legs may import this module.
"""
import multiprocessing as mp
import os

import numpy as np

from . import krotov_oracle as orc


def _worker(conn, terms, psi0, L, tlist, is_super):
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=1)
    except Exception:   # pragma: no cover
        pass
    Kb = len(terms)
    nt = len(tlist)
    adj = [orc.adjoint_terms(t) for t in terms]
    mus = [[orc.mu_operator(terms[k], l, is_super) for l in range(L)]
           for k in range(Kb)]
    while True:
        msg = conn.recv()
        if msg[0] == 'stop':
            return
        if msg[0] == 'fw0':      # optimize.py:806-846, no communication
            pulses = msg[1]
            conn.send([orc.forward_propagation(terms[k], pulses, tlist,
                                               psi0[k], is_super,
                                               store_all=False)
                       for k in range(Kb)])
            continue
        # one Krotov iteration: ('iter', chis_block, chi_norms_block, guess)
        _, chis, chi_norms, guess = msg
        X = [orc.backward_propagation(adj[k], guess, tlist, chis[k], is_super)
             for k in range(Kb)]
        optimized = [p.copy() for p in guess]
        fw = [np.asarray(p, dtype=np.complex128) for p in psi0]
        for n in range(nt - 1):
            dt = tlist[n + 1] - tlist[n]
            part = np.zeros(L, dtype=np.complex128)
            for l in range(L):
                acc = 0j
                for k in range(Kb):
                    if mus[k][l] is None:
                        continue
                    acc += complex(np.vdot(X[k][n], mus[k][l] @ fw[k])) \
                        * chi_norms[k]
                part[l] = acc
            conn.send_bytes(part.tobytes())
            new = np.frombuffer(conn.recv_bytes(), dtype=np.float64)
            for l in range(L):
                optimized[l][n] = new[l]
            fw = [orc.expm_step(terms[k], optimized, n, dt, fw[k], is_super,
                                False) for k in range(Kb)]
        conn.send(fw)


class ParallelOracle:
    """First-order Krotov iterations of the numpy oracle on `nproc` forked
    worker processes (default: all host CPUs, at most one per objective)."""

    def __init__(self, terms, psi0, targets, shapes, lambdas, tlist,
                 is_super=False, nproc=None):
        K = len(terms)
        nproc = max(1, min(K, nproc or os.cpu_count() or 1))
        self.K, self.L, self.nproc = K, len(shapes), nproc
        self.targets, self.shapes, self.lambdas = targets, shapes, lambdas
        self.tlist, self.is_super = np.asarray(tlist), is_super
        bounds = np.linspace(0, K, nproc + 1).astype(int)
        self.blocks = [(int(a), int(b)) for a, b in zip(bounds, bounds[1:])]
        ctx = mp.get_context('fork')
        self.conns, self.procs = [], []
        for a, b in self.blocks:
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, daemon=True,
                            args=(child, terms[a:b], psi0[a:b], self.L,
                                  self.tlist, is_super))
            p.start()
            child.close()
            self.conns.append(parent)
            self.procs.append(p)

    def forward(self, pulses):
        for c in self.conns:
            c.send(('fw0', pulses))
        out = []
        for c in self.conns:
            out.extend(c.recv())
        return out

    def iteration(self, guess_pulses, fw_states_T, tau_vals,
                  chi_constructor=orc.chis_re, weights=None):
        """optimize.py:393-508 with the three maps of
        parallelization.py:51-57."""
        chis = chi_constructor(fw_states_T, self.targets, tau_vals, weights)
        norms = [orc.state_norm(c, self.is_super) for c in chis]
        chis = [c / nrm for c, nrm in zip(chis, norms)]
        for c, (a, b) in zip(self.conns, self.blocks):
            c.send(('iter', chis[a:b], norms[a:b], guess_pulses))
        L, tlist = self.L, self.tlist
        optimized = [np.array(p, dtype=np.float64) for p in guess_pulses]
        g_a = np.zeros(L)
        for n in range(len(tlist) - 1):
            dt = tlist[n + 1] - tlist[n]
            acc = np.zeros(L, dtype=np.complex128)
            for c in self.conns:        # fixed worker order: deterministic
                acc += np.frombuffer(c.recv_bytes(), dtype=np.complex128)
            new = np.empty(L)
            for l in range(L):          # optimize.py:471-477
                S_t = self.shapes[l][n]
                d1 = acc[l].imag
                g_a[l] += (S_t / self.lambdas[l]) * abs(d1) ** 2 * dt
                optimized[l][n] += (S_t / self.lambdas[l]) * d1
                new[l] = optimized[l][n]
            buf = new.tobytes()
            for c in self.conns:
                c.send_bytes(buf)
        fw = []
        for c in self.conns:
            fw.extend(c.recv())
        tau = np.array([complex(np.vdot(self.targets[k], fw[k]))
                        for k in range(self.K)])
        return dict(optimized_pulses=optimized, g_a=g_a, tau_vals=tau,
                    fw_states_T=fw)

    def close(self):
        for c in self.conns:
            try:
                c.send(('stop',))
            except Exception:   # pragma: no cover
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():    # pragma: no cover
                p.terminate()
        self.conns, self.procs = [], []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
