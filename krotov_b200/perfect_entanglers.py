"""Perfect-entangler functional for two-qubit gates and its chi constructor.

The reference does not contain this functional: its notebook 07
(docs/notebooks/07_example_PE.ipynb, cells 25-39) takes
``make_PE_krotov_chi_constructor`` and ``F_PE`` from the third-party package
``weylchamber`` (0.4.0, not vendored).  This module restates the published
algorithm -- Watts et al., Phys. Rev. A 91, 062306 (2015); Goerz et al., Phys.
Rev. A 91, 062307 (2015), Eq. (33b); the local invariants of Makhlin, Quantum
Inf. Process. 1, 243 (2002) -- with the names a notebook uses, so that
``gate_objectives(basis, 'PE', H)`` has a matching ``chi_constructor``:

* ``gate(basis, states)``     U_ij = <basis_i | states_j>
* ``g1g2g3(U)``               local invariants of a two-qubit gate
* ``F_PE(g1, g2, g3)``        g3 sqrt(g1^2 + g2^2) - g1  (<= 0 for perfect entanglers)
* ``make_PE_krotov_chi_constructor(canonical_basis)``
                              chi_l = -dF_PE / d<phi_l(T)| for the four
                              forward-propagated Bell states phi_l(T)

With a_kl = <B_k|phi_l(T)> in the Bell ("magic") basis of
``gate_objectives(..., 'PE')`` (objectives.py:1042-1047 of the reference) and
m = a^T a, t = tr m, D = det a, the invariants are the holomorphic functions
G = t^2 / (16 D) = g1 + i g2 and g3 = Re[(t^2 - tr m^2) / (4 D)], so that

    dF/da_kl = (g3 conj(G) / (2|G|) - 1/2) dG/da_kl + (|G| / 2) dg3c/da_kl,
    dG/da    = t a / (4 D) - G a^{-T},
    dg3c/da  = (t a - a a^T a) / D - g3c a^{-T},

and dF/da*_kl = conj(dF/da_kl) because F is real.  The chi states are host
arrays (one 4 x 4 gate per Krotov iteration): the engine uploads them like the
result of any other user ``chi_constructor``.

PARITY: unpinned against ``weylchamber`` itself (absent here, no network).
Pinned: the local invariants of 1, CNOT and SWAP; F_PE = 1.447335 of the guess
in notebook 07 (cell 39, iteration 0); the gradient against central
differences (1e-10); an optimisation of notebook 07's problem reaches a perfect
entangler (F_PE < 0) within the notebook's 8 iterations.  The invariants are
divided by det(a), i.e. they are the U(4) form (any global phase); on gates
with det = 1 they equal the SU(4) form tr(m)^2/16 that weylchamber may use, but
the two have different derivatives NORMAL to the unitary group, which the
non-perturbative Krotov update sees: the iterates after iteration 0 therefore
need not reproduce notebook 07's printed F_PE values digit by digit (ours:
1.0768, 0.6504, ... against 0.9981, 0.5826, ...).
"""
import numpy as np

from ._dense import dense

__all__ = ['gate', 'to_magic', 'from_magic', 'g1g2g3', 'F_PE', 'F_PE_gradient',
           'make_PE_krotov_chi_constructor']

# columns = the Bell states of objectives.py:1044-1047 in the canonical basis
MAGIC = np.array([[1, 0, 0, 1j],
                  [0, 1j, 1, 0],
                  [0, 1j, -1, 0],
                  [1, 0, 0, -1j]], dtype=np.complex128) / np.sqrt(2.0)


def _vec(state):
    return np.asarray(dense(state), dtype=np.complex128).reshape(-1)


def gate(basis, states):
    """U_ij = <basis_i | states_j> (``weylchamber.gates.gate``)."""
    B = np.array([_vec(b) for b in basis])
    S = np.array([_vec(s) for s in states])
    return B.conj() @ S.T


def to_magic(U):
    """Q^dag U Q: the gate in the Bell basis."""
    return MAGIC.conj().T @ np.asarray(U, dtype=np.complex128) @ MAGIC


def from_magic(UB):
    """Q UB Q^dag: a gate given in the Bell basis, back in the canonical one."""
    return MAGIC @ np.asarray(UB, dtype=np.complex128) @ MAGIC.conj().T


def _invariants(a):
    m = a.T @ a
    t, t2, D = np.trace(m), np.trace(m @ m), np.linalg.det(a)
    return m, t, t2, D, t * t / (16.0 * D), (t * t - t2) / (4.0 * D)


def g1g2g3(U):
    """Local invariants (g1, g2, g3) of the two-qubit gate U (canonical
    basis)."""
    _, _, _, _, G, g3c = _invariants(to_magic(U))
    return G.real, G.imag, g3c.real


def F_PE(g1, g2, g3):
    """Perfect-entangler functional; zero on the surface of the polyhedron of
    perfect entanglers, negative inside."""
    return g3 * np.sqrt(g1 * g1 + g2 * g2) - g1


def F_PE_gradient(a):
    """(F_PE, dF_PE/da*) for the gate `a` given in the Bell basis."""
    a = np.asarray(a, dtype=np.complex128)
    m, t, t2, D, G, g3c = _invariants(a)
    ainvT = np.linalg.inv(a).T
    dG = t * a / (4.0 * D) - G * ainvT
    dg3 = (t * a - a @ m) / D - g3c * ainvT
    absG = abs(G)
    wG = (g3c.real * np.conj(G) / (2.0 * absG) - 0.5) if absG > 0 else -0.5
    dF = wG * dG + 0.5 * absG * dg3
    return F_PE(G.real, G.imag, g3c.real), np.conj(dF)


def make_PE_krotov_chi_constructor(canonical_basis, unitarity_weight=0):
    """chi_constructor for an optimisation towards a perfect entangler
    (``weylchamber.perfect_entanglers.make_PE_krotov_chi_constructor``):
    ``chi_l = -dF_PE/d<phi_l(T)|`` for the forward-propagated Bell states of
    ``gate_objectives(canonical_basis, 'PE', H)``, in the order of the
    objectives."""
    if unitarity_weight != 0:
        raise NotImplementedError(
            "unitarity_weight != 0 (loss from the logical subspace) is not "
            "restated")
    basis = np.array([_vec(b) for b in canonical_basis])      # rows <- kets
    bell = MAGIC.T @ basis                                     # rows = Bell kets

    def chi_constructor(fw_states_T, *args, **kwargs):
        from ._dense import like
        phi = np.array([_vec(s) for s in fw_states_T])         # rows = phi_l
        a = bell.conj() @ phi.T                                # a_kl
        _, dFc = F_PE_gradient(a)
        chis = []
        for l, template in enumerate(fw_states_T):
            v = -(dFc[:, l] @ bell)
            chis.append(like(template, v.reshape(np.shape(dense(template)))))
        return chis

    return chi_constructor
