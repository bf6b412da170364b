"""Single-time-step propagators with the reference's plugin signature
``propagator(H, state, dt, c_ops=None, backwards=False, initialize=False)``
(/root/reference/src/krotov/propagators.py:13-47).

:func:`expm` and :class:`DensityMatrixODEPropagator` are *markers* for
:func:`krotov_b200.optimize_pulses`: when one of them is passed as
`propagator`, the whole time loop is lowered to the sm_100a sweep kernels
(exact piecewise-constant propagation ``exp(f A dt) v`` evaluated as a scaled
Taylor series on the device, see csrc/kq_common.cuh) and the functions below
are never called per step.  They remain callable on host objects for
step-wise use such as :meth:`Objective.propagate`; that host path is an
analysis helper, not a fallback for the optimisation.
"""
from abc import ABC, abstractmethod

import numpy as np
import scipy.linalg

from ._dense import dense, kind_of, like

__all__ = ['expm', 'Propagator', 'DensityMatrixODEPropagator']


def _assemble(H, factor):
    A = None
    for part in (H if isinstance(H, list) else [H]):
        if isinstance(part, list):
            term = (factor * part[1]) * dense(part[0])
        else:
            term = factor * dense(part)
        A = term if A is None else A + term
    return A


def expm(H, state, dt, c_ops=None, backwards=False, initialize=False):
    """Propagate `state` by ``exp(f*A*dt)`` with ``A = H0 + sum c_m H_m``;
    f = -i in Hilbert space, 1 for a super-operator acting on a density
    matrix (column-stacking), conjugated for ``backwards``
    (propagators.py:79-122).  Collapse operators are not supported."""
    if c_ops is None:
        c_ops = []
    if len(c_ops) > 0:
        raise NotImplementedError("Liouville exponentiation not implemented")
    assert isinstance(H, list) and len(H) > 0
    first = H[0][0] if isinstance(H[0], list) else H[0]
    a0, s = dense(first), dense(state)
    is_ket = s.shape[1] == 1 and a0.shape[0] == s.shape[0]
    is_super = (not is_ket) and a0.shape[0] == s.shape[0] * s.shape[1]
    if not (is_ket or is_super) or (kind_of(first) == 'super' and is_ket):
        raise NotImplementedError(
            "Cannot handle argument types A:%s, state:%s"
            % (kind_of(first), kind_of(state))
        )
    factor = 1 if is_super else -1j
    if backwards:
        factor = np.conjugate(factor)
    U = scipy.linalg.expm(_assemble(H, factor) * dt)
    if is_super:
        d = s.shape[0]
        out = (U @ s.reshape(-1, order='F')).reshape(d, d, order='F')
    else:
        out = U @ s
    return like(state, out)


class Propagator(ABC):
    """Base class for stateful propagators (propagators.py:125-159)."""

    @abstractmethod
    def __call__(self, H, state, dt, c_ops=None, backwards=False,
                 initialize=False):
        """Propagate `state` over one time step `dt`."""


class DensityMatrixODEPropagator(Propagator):
    """Density-matrix propagator for a Liouvillian in nested-list form.

    The reference integrates with zvode (Adams, adaptive) and carries the
    multistep history across time steps (propagators.py:162-327), which makes
    its output depend on the integrator's internal step sequence (its own
    notebook marks the result as not reproducible across systems).  On the
    B200 engine the piecewise-constant Liouvillian is propagated *exactly*,
    ``rho <- vec^-1(exp(L dt) vec rho)``, by the same sweep kernels as
    :func:`expm`; the tolerance arguments are accepted for interface
    compatibility and ignored.  See DESIGN.md for the measured deviation
    between the two (1e-3 relative in the pulse at nt=2500).
    """

    def __init__(self, method='adams', order=12, atol=1e-8, rtol=1e-6,
                 nsteps=1000, first_step=0, min_step=0, max_step=0,
                 reentrant=False):
        self.method, self.order = method, order
        self.atol, self.rtol, self.nsteps = atol, rtol, nsteps
        self.first_step, self.min_step, self.max_step = (
            first_step, min_step, max_step)
        self.reentrant = reentrant

    def __call__(self, H, state, dt, c_ops=None, backwards=False,
                 initialize=False):
        if not (c_ops is None or len(c_ops) == 0):
            raise NotImplementedError("c_ops not implemented")
        # `backwards` has no effect in Liouville space (propagators.py:231-235)
        return expm(H, state, dt, None, backwards=False)
