"""Final-time functionals and their chi-constructors.

Names, arguments and values follow ``krotov.functionals``
(/root/reference/src/krotov/functionals.py).  The four state-based
constructors (:func:`chis_re`, :func:`chis_ss`, :func:`chis_sm`,
:func:`chis_hs`) are recognised by :func:`krotov_b200.optimize_pulses` and
evaluated on the device (``kq_chi_boundary``), so that an iteration needs no
host round trip; any other callable with the ``chi_constructor`` signature
runs as a host callback once per iteration.  The host implementations below
work on ndarray or Qobj-like states.
"""
import logging

import numpy as np

from ._dense import dense
from .second_order import _overlap

__all__ = [
    'f_tau', 'F_ss', 'J_T_ss', 'chis_ss', 'F_sm', 'J_T_sm', 'chis_sm',
    'F_re', 'J_T_re', 'chis_re', 'J_T_hs', 'chis_hs', 'F_avg', 'gate',
    'mapped_basis',
]


def _taus(fw_states_T, objectives, tau_vals):
    if tau_vals is None:
        tau_vals = [_overlap(obj.target, psi)
                    for psi, obj in zip(fw_states_T, objectives)]
    return tau_vals


def _weight(obj):
    return getattr(obj, 'weight', None)


def f_tau(fw_states_T, objectives, tau_vals=None, **kwargs):
    r""":math:`f_\tau = \frac1N\sum_k w_k\tau_k` (functionals.py:82-132)."""
    res = 0j
    for obj, tau in zip(objectives, _taus(fw_states_T, objectives, tau_vals)):
        if tau is None:
            logging.getLogger('krotov').warning("τ is None in f_tau")
            continue
        w = _weight(obj)
        res += tau if w is None else w * tau
    return res / len(objectives)


def F_ss(fw_states_T, objectives, tau_vals=None, **kwargs):
    r""":math:`\frac1N\sum_k w_k|\tau_k|^2`."""
    taus = _taus(fw_states_T, objectives, tau_vals)
    F = f_tau(fw_states_T, objectives, [abs(t) ** 2 for t in taus])
    assert abs(F.imag) < 1e-10, F.imag
    return F.real


def J_T_ss(fw_states_T, objectives, tau_vals=None, **kwargs):
    return 1 - F_ss(fw_states_T, objectives, tau_vals)


def chis_ss(fw_states_T, objectives, tau_vals):
    r""":math:`\chi_k=\frac1N w_k\tau_k\,\Psi_k^{tgt}` (functionals.py:177-197)."""
    N = len(objectives)
    out = []
    for obj, tau in zip(objectives, tau_vals):
        w = _weight(obj)
        c = (tau / N) if w is None else (tau / N) * w
        out.append(c * obj.target)
    return out


def F_sm(fw_states_T, objectives, tau_vals=None, **kwargs):
    return abs(f_tau(fw_states_T, objectives, tau_vals)) ** 2


def J_T_sm(fw_states_T, objectives, tau_vals=None, **kwargs):
    return 1 - F_sm(fw_states_T, objectives, tau_vals)


def chis_sm(fw_states_T, objectives, tau_vals):
    r""":math:`\chi_k=\frac{1}{N^2}w_k\sum_j w_j\tau_j\,\Psi_k^{tgt}`
    (functionals.py:225-253)."""
    s = 0
    for obj, tau in zip(objectives, tau_vals):
        w = _weight(obj)
        s += tau if w is None else w * tau
    c = 1.0 / len(objectives) ** 2
    out = []
    for obj in objectives:
        w = _weight(obj)
        out.append(c * obj.target * s if w is None
                   else c * w * obj.target * s)
    return out


def F_re(fw_states_T, objectives, tau_vals=None, **kwargs):
    return f_tau(fw_states_T, objectives, tau_vals).real


def J_T_re(fw_states_T, objectives, tau_vals=None, **kwargs):
    return 1 - F_re(fw_states_T, objectives, tau_vals)


def chis_re(fw_states_T, objectives, tau_vals):
    r""":math:`\chi_k=\frac{1}{2N}w_k\,\Psi_k^{tgt}` (functionals.py:293-317)."""
    c = 1.0 / (2 * len(objectives))
    out = []
    for obj in objectives:
        w = _weight(obj)
        out.append(c * obj.target if w is None else c * w * obj.target)
    return out


def J_T_hs(fw_states_T, objectives, tau_vals=None, **kwargs):
    r""":math:`\frac{1}{2N}\sum_k w_k\|\rho_k(T)-\rho_k^{tgt}\|^2_{hs}`
    (functionals.py:320-386)."""
    res = 0.0
    for obj, psi in zip(objectives, fw_states_T):
        d = dense(psi) - dense(obj.target)
        w = _weight(obj)
        res += (1.0 if w is None else w) * float(np.vdot(d, d).real)
    return res / (2 * len(objectives))


def chis_hs(fw_states_T, objectives, tau_vals):
    r""":math:`\chi_k=\frac{1}{2N}w_k(\rho_k^{tgt}-\rho_k(T))`
    (functionals.py:389-437)."""
    c = 1.0 / (2 * len(objectives))
    out = []
    for obj, psi in zip(objectives, fw_states_T):
        w = _weight(obj)
        out.append(c * (obj.target - psi) if w is None
                   else c * w * (obj.target - psi))
    return out


def F_avg(fw_states_T, basis_states, gate, mapped_basis_states=None,
          prec=1e-5):
    r"""Average gate fidelity from the N propagated logical basis kets
    (Hilbert space: :math:`(|\mathrm{tr}\,O^\dagger U|^2 +
    \mathrm{tr}\,O^\dagger U U^\dagger O)/(N(N+1))`) or from the N^2
    propagated basis dyads :math:`\rho_{ij}` (Liouville space),
    functionals.py:440-570 of the reference.

    Raises:
        ValueError: wrong number of states for the space.
    """
    N = len(basis_states)
    O = np.asarray(gate.full() if hasattr(gate, 'full') else gate,
                   dtype=np.complex128)
    first = dense(fw_states_T[0])
    if first.shape[1] == 1:
        if len(fw_states_T) != N:
            raise ValueError(
                "Evaluating F_avg for hilbert space states requires %d states "
                "(forward-propagation of all basis states), not %d"
                % (N, len(fw_states_T)))
        U = globals()['gate'](basis_states, fw_states_T)
        OU = O.conj().T @ U
        F = abs(np.trace(OU)) ** 2 + np.trace(OU @ OU.conj().T)
    else:
        if len(fw_states_T) != N * N:
            raise ValueError(
                "Evaluating F_avg for density matrices requires %d states "
                "(forward-propagation of all dyadic combinations of "
                "%d basis states), not %d" % (N * N, N, len(fw_states_T)))
        if mapped_basis_states is None:
            mapped_basis_states = mapped_basis(O, basis_states)
        mapped = [dense(m) for m in mapped_basis_states]
        F = 0
        for j in range(N):
            rho_jj = dense(fw_states_T[j * N + j])
            for i in range(N):
                rho_ij = dense(fw_states_T[i * N + j])
                F += np.vdot(mapped[i], rho_ij @ mapped[j])
                F += np.vdot(mapped[i], rho_jj @ mapped[i])
    assert abs(complex(F).imag) < prec, "%.2e > %.2e" % (
        complex(F).imag, prec)
    return float(complex(F).real) / (N * (N + 1))


def gate(basis_states, fw_states_T):
    """Matrix ``U[i,j] = <basis_i|fw_state_j>`` of the implemented gate."""
    N = len(basis_states)
    U = np.zeros((N, N), dtype=np.complex128)
    for j in range(N):
        for i in range(N):
            U[i, j] = np.vdot(dense(basis_states[i]), dense(fw_states_T[j]))
    return U


def mapped_basis(O, basis_states):
    """States ``sum_i O[i,j] basis_i`` for every column j."""
    O = np.asarray(O.full() if hasattr(O, 'full') else O)
    return tuple(
        sum(complex(O[i, j]) * basis_states[i] for i in range(O.shape[0]))
        for j in range(O.shape[1])
    )
