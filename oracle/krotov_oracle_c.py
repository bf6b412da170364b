"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the C oracle
(oracle/krotov_oracle_c.c, built by oracle/Makefile into oracle/_ref/).

Same algorithm as oracle/krotov_oracle.py with its own Pade matrix
exponential and OpenMP over the objectives; used for the multi-threaded CPU
baseline in bench.py and for fast checks of large configurations.  It is
pinned against the numpy oracle (and through it against the reference's
golden vectors) in tests/test_oracle.py.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_ref', 'libkrotov_oracle.so')
_lib = None


def build():
    subprocess.run(['make', '-C', HERE], check=True,
                   stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def max_threads():
    return int(load().kqo_max_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class COracle:
    """Dense problem (same description as Workload.lowered()) on the C side."""

    def __init__(self, low, nthreads=1):
        self.lib = load()
        self.nthreads = int(nthreads)
        terms = low['terms']
        self.K = len(terms)
        self.N = len(low['psi0'][0])
        self.L = len(low['pulses'])
        self.NT = len(low['tlist']) - 1
        self.M = max(len(t) for t in terms)
        K, N, L, M = self.K, self.N, self.L, self.M
        self.ops = np.zeros((K, M, N, N), dtype=np.complex128)
        self.ops_adj = np.zeros((K, M, N, N), dtype=np.complex128)
        self.t2p = np.full((K, M), -2, dtype=np.int32)
        self.mu = np.zeros((K, max(L, 1), N, N), dtype=np.complex128)
        self.is_super = 1 if low['is_super'] else 0
        for k, tk in enumerate(terms):
            for m, (op, l) in enumerate(tk):
                self.ops[k, m] = op
                self.ops_adj[k, m] = op.conj().T
                self.t2p[k, m] = l
                if l >= 0:
                    self.mu[k, l] += (1j * op) if self.is_super else op
        tl = np.asarray(low['tlist'], dtype=np.float64)
        self.dt = np.array([tl[n + 1] - tl[n] for n in range(self.NT)])
        self.shape = np.ascontiguousarray(np.array(low['shapes'], dtype=np.float64))
        self.lam = np.ascontiguousarray(np.array(low['lambdas'], dtype=np.float64))
        self.psi0 = np.ascontiguousarray(np.array(low['psi0'], dtype=np.complex128))
        self.targets = np.ascontiguousarray(
            np.array(low['targets'], dtype=np.complex128))

    def forward(self, pulses):
        pulses = np.ascontiguousarray(np.array(pulses, dtype=np.float64))
        phiT = np.zeros((self.K, self.N), dtype=np.complex128)
        rc = self.lib.kqo_forward(
            self.K, self.N, self.NT, self.L, self.M, _p(self.ops), _p(self.t2p),
            _p(self.dt), _p(pulses), _p(self.psi0), _p(phiT), None,
            self.is_super, self.nthreads)
        assert rc == 0
        return phiT

    def iteration(self, pulses, chis):
        """One Krotov iteration from the (un-normalised) boundary states."""
        pulses = np.ascontiguousarray(np.array(pulses, dtype=np.float64))
        chis = np.array(chis, dtype=np.complex128).reshape(self.K, self.N)
        norms = np.ascontiguousarray(np.linalg.norm(chis, axis=1))
        chiT = np.ascontiguousarray(chis / norms[:, None])
        X = np.zeros((self.K, self.NT + 1, self.N), dtype=np.complex128)
        rc = self.lib.kqo_backward(
            self.K, self.N, self.NT, self.L, self.M, _p(self.ops_adj),
            _p(self.t2p), _p(self.dt), _p(pulses), _p(chiT), _p(X),
            self.is_super, self.nthreads)
        assert rc == 0
        opt = np.zeros_like(pulses)
        phiT = np.zeros((self.K, self.N), dtype=np.complex128)
        g_a = np.zeros(self.L)
        rc = self.lib.kqo_update_sweep(
            self.K, self.N, self.NT, self.L, self.M, _p(self.ops), _p(self.mu),
            _p(self.t2p), _p(self.dt), _p(self.shape), _p(self.lam),
            _p(pulses), _p(opt), _p(X), _p(norms), _p(self.psi0), _p(phiT),
            _p(g_a), self.is_super, self.nthreads)
        assert rc == 0
        return dict(optimized_pulses=opt, g_a=g_a, fw_states_T=phiT,
                    backward_states=X,
                    tau_vals=np.einsum('kn,kn->k', self.targets.conj(), phiT))

    def optimize(self, pulses, chi_constructor, iters, weights=None):
        """chi_constructor as in krotov_oracle (fw_T, targets, tau, weights)."""
        pulses = np.array(pulses, dtype=np.float64)
        phiT = self.forward(pulses)
        tau = np.einsum('kn,kn->k', self.targets.conj(), phiT)
        records = []
        for _ in range(iters):
            chis = chi_constructor(list(phiT), list(self.targets), list(tau),
                                   weights)
            rec = self.iteration(pulses, chis)
            records.append(rec)
            pulses, phiT, tau = rec['optimized_pulses'], rec['fw_states_T'], \
                rec['tau_vals']
        return records
