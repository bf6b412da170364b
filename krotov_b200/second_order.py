"""Second-order Krotov support: the ``Sigma`` interface and the numerical
estimate of its parameter A (reference: src/krotov/second_order.py).

In the B200 engine sigma(t) is evaluated on the host at all interval
midpoints once per iteration (it only changes in :meth:`Sigma.refresh`,
optimize.py:566-577) and uploaded as a float64 array; the term
``0.5*sigma*<dphi_k| mu |phi_k>`` (optimize.py:468-469) is fused into the
forward/update sweep kernel.
"""
from abc import ABC, abstractmethod

import numpy as np

from ._dense import dense, kind_of

__all__ = ['Sigma', 'numerical_estimate_A']


class Sigma(ABC):
    """Function sigma(t) of the second-order update; subclass and implement
    :meth:`__call__` and :meth:`refresh` (second_order.py:9-66)."""

    @abstractmethod
    def __call__(self, t):
        """Value of sigma at time `t` (real)."""

    @abstractmethod
    def refresh(self, forward_states, forward_states0, chi_states, chi_norms,
                optimized_pulses, guess_pulses, objectives, result):
        """Recalculate internal parameters after an iteration; arguments as
        in the reference (second_order.py:34-66).  `forward_states` and
        `forward_states0` are per-objective sequences indexable by time
        index, backed by the device stores and downloaded on first use."""


def _overlap(a, b):
    """<a|b> for kets, tr(a^dag b) for operators; None for non-states
    (second_order.py:69-83)."""
    try:
        if a is None or b is None or isinstance(a, str) or isinstance(b, str):
            return None
        ka, kb = kind_of(a), kind_of(b)
        A, B = dense(a), dense(b)
        if A.shape != B.shape and not (ka == 'bra' or kb == 'bra'):
            return None
        if ka == 'bra':
            A = A.conj().T
        if kb == 'bra':
            B = B.conj().T
        return complex(np.vdot(A.reshape(-1), B.reshape(-1)))
    except (AttributeError, TypeError, ValueError):
        return None


def numerical_estimate_A(forward_states, forward_states0, chi_states,
                         chi_norms, Delta_J_T):
    r"""Estimate :math:`A = (\sum_k 2\Re\langle\chi_k(T)|\Delta\phi_k(T)\rangle
    + \Delta J_T) / \sum_k \|\Delta\phi_k(T)\|^2` (second_order.py:86-141);
    returns 0 when the states did not change."""
    n = len(forward_states0)
    dphi = [dense(forward_states[k][-1]) - dense(forward_states0[k][-1])
            for k in range(n)]
    denom = sum(np.vdot(d.reshape(-1), d.reshape(-1)).real for d in dphi)
    if denom > 1.0e-30:
        numer = sum(
            (2 * chi_norms[k] * np.vdot(dense(chi_states[k]).reshape(-1),
                                        dphi[k].reshape(-1))).real
            for k in range(n)
        ) + Delta_J_T
        return numer / denom
    return 0
