"""Synthetic problem definitions for the five BASELINE.json ``configs``
(labelled C1..C5 as in SURVEY.md §8) plus the reference's own test fixtures.

Every workload is described with plain numpy operators and callable controls
in QuTiP's nested-list format, so the *same* description can be handed to

* :func:`krotov_b200.optimize_pulses` (the CUDA path),
* the unmodified reference package (``oracle/make_golden.py``, this
  container only), either as numpy objects or wrapped into Qobj,
* the numpy oracle (``oracle/krotov_oracle.py``) through
  :meth:`Workload.lowered`.

Sources of the physics (all under /root/reference):
C1 ``docs/notebooks/01_example_simple_state_to_state.ipynb`` /
``tests/test_krotov.py:137-163``; C2 ``tests/transmon_xgate_system_mod.py``;
C3 ``docs/notebooks/07_example_PE.ipynb`` cell 12; C4 detuned ensemble of C1
systems (cf. ``docs/notebooks/08_example_ensemble.ipynb``); C5
``docs/notebooks/04_example_dissipative_qubit_reset.ipynb`` cells 5-24.
"""
from dataclasses import dataclass, field
from functools import partial

import numpy as np

from . import shapes
from .conversions import (control_onto_interval, discretize, extract_controls,
                          pulse_options_dict_to_list)

__all__ = [
    'Workload', 'tls_state_to_state', 'tls_reference_fixture',
    'transmon_xgate', 'two_qubit_gate', 'tls_ensemble',
    'dissipative_qubit_reset', 'lambda_system', 'tls_shared_controls',
    'by_name',
]

_SX = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_SY = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_SZ = np.array([[1, 0], [0, -1]], dtype=np.complex128)


def _ket(n, i):
    v = np.zeros((n, 1), dtype=np.complex128)
    v[i, 0] = 1
    return v


@dataclass
class Workload:
    """A complete optimisation problem in nested-list / numpy form."""

    name: str
    Hs: list                # per objective: [H0, [H1, control], ...]
    initial_states: list    # per objective: (N,1) ket or (d,d) density matrix
    targets: list           # per objective
    pulse_options: dict     # control -> dict(lambda_a=..., update_shape=...)
    tlist: np.ndarray
    chi: str                # 're' | 'ss' | 'sm' | 'hs' | 'qubit_reset'
    is_super: bool = False
    weights: list = None
    meta: dict = field(default_factory=dict)

    @property
    def K(self):
        return len(self.Hs)

    @property
    def nt(self):
        return len(self.tlist)

    def objectives(self, Objective, wrap=None):
        """Build objectives with the given ``Objective`` class; `wrap` maps
        every numpy operator/state to the caller's object type (e.g. Qobj)."""
        w = (lambda a: a) if wrap is None else wrap
        cache = {}

        def wrap_once(a):
            key = id(a)
            if key not in cache:
                cache[key] = w(a)
            return cache[key]

        out = []
        for k in range(self.K):
            H = [
                [wrap_once(h[0]), h[1]] if isinstance(h, list) else wrap_once(h)
                for h in self.Hs[k]
            ]
            tgt = self.targets[k]
            obj = Objective(
                initial_state=wrap_once(self.initial_states[k]),
                target=wrap_once(tgt) if isinstance(tgt, np.ndarray) else tgt,
                H=H,
            )
            if self.weights is not None:
                obj.weight = self.weights[k]
            out.append(obj)
        return out

    def lowered(self):
        """Dense description for the oracle: ``terms[k] = [(op, l), ...]``,
        vectorised states/targets (column-stacking for density matrices),
        interval pulses ``[L][nt-1]``, shapes, lambdas."""
        class _O:  # minimal objective-like for extract_controls
            def __init__(self, H):
                self.H, self.c_ops = H, []
        controls = extract_controls([_O(H) for H in self.Hs])
        opts = pulse_options_dict_to_list(self.pulse_options, controls)
        pulses, shp, lam = [], [], []
        for c, o in zip(controls, opts):
            ctl = discretize(c, self.tlist, args=(o.get('args', None),),
                             via_midpoints=True)
            pulses.append(control_onto_interval(ctl))
            S = o['update_shape']
            S = {1: shapes.one_shape, 0: shapes.zero_shape}.get(S, S) \
                if not callable(S) else S
            s_arr = control_onto_interval(
                discretize(S, self.tlist, args=(), via_midpoints=True))
            shp.append(np.clip(s_arr, 0.0, 1.0))
            lam.append(float(o['lambda_a']))

        def idx(ctrl):
            for i, c in enumerate(controls):
                if c is ctrl:
                    return i
            raise KeyError

        terms = [
            [(np.asarray(h[0], dtype=np.complex128), idx(h[1]))
             if isinstance(h, list)
             else (np.asarray(h, dtype=np.complex128), -1) for h in H]
            for H in self.Hs
        ]

        def vec(s):
            s = np.asarray(s, dtype=np.complex128)
            if s.ndim == 2 and s.shape[1] > 1:
                return s.reshape(-1, order='F')
            return s.reshape(-1)

        return dict(
            terms=terms,
            psi0=[vec(s) for s in self.initial_states],
            targets=[vec(t) for t in self.targets],
            pulses=pulses, shapes=shp, lambdas=np.array(lam),
            tlist=self.tlist, is_super=self.is_super, weights=self.weights,
        )


# --- C1 ---------------------------------------------------------------------

def tls_state_to_state(nt=500, T=5.0, omega=1.0, ampl0=0.2, lambda_a=5.0):
    """C1: two-level |0> -> |1>, Blackman-shaped guess, ``chis_ss``
    (notebook 01 cells 5-13, 25-28)."""
    H0 = -0.5 * omega * _SZ
    guess = lambda t, args: ampl0 * shapes.flattop(  # noqa: E731
        t, t_start=0, t_stop=T, t_rise=0.3, func='blackman')
    S = partial(shapes.flattop, t_start=0, t_stop=T, t_rise=0.3, t_fall=0.3,
                func='blackman')
    return Workload(
        name='C1_tls_state_to_state',
        Hs=[[H0, [_SX.copy(), guess]]],
        initial_states=[_ket(2, 0)], targets=[_ket(2, 1)],
        pulse_options={guess: dict(lambda_a=lambda_a, update_shape=S)},
        tlist=np.linspace(0, T, nt), chi='ss',
    )


def tls_reference_fixture(nt=500):
    """The reference's own integration-test system
    (tests/test_krotov.py:137-163): constant guess 0.2, sin² flat-top shape,
    lambda_a=5, ``chis_re``; golden table tests/test_krotov/oct.log."""
    H0 = -0.5 * _SZ
    guess = lambda t, args: 0.2  # noqa: E731
    S = partial(shapes.flattop, t_start=0, t_stop=5, t_rise=0.3, t_fall=0.3,
                func='sinsq')
    return Workload(
        name='tls_reference_fixture',
        Hs=[[H0, [_SX.copy(), guess]]],
        initial_states=[_ket(2, 0)], targets=[_ket(2, 1)],
        pulse_options={guess: dict(lambda_a=5, update_shape=S)},
        tlist=np.linspace(0, 5, nt), chi='re',
    )


# --- C2 ---------------------------------------------------------------------

def _transmon_guess(t, args, T=10.0):
    return 4 * np.exp(-40.0 * (t / T - 0.5) ** 2)


def transmon_xgate(nstates=1, nt=1000, T=10.0, Ec=0.386, EjEc=45, ng=0.0,
                   lambda_a=1.0, guess=None):
    """C2: transmon X gate, two objectives from ``gate_objectives(sigma_x)``
    on the two lowest eigenstates (tests/transmon_xgate_system_mod.py:14-44,
    tests/test_parallelization.py:63-110).  ``nstates=1`` gives N=3 (C2),
    ``2`` gives the reference test's N=5."""
    Ej = EjEc * Ec
    n = np.arange(-nstates, nstates + 1)
    up = np.diag(np.ones(2 * nstates), k=-1)
    H0 = (np.diag(4 * Ec * (n - ng) ** 2) - Ej * (up + up.T) / 2.0).astype(
        np.complex128)
    H1 = (-2 * np.diag(n)).astype(np.complex128)
    evals, evecs = np.linalg.eigh(H0.real)
    psi0 = evecs[:, 0].reshape(-1, 1).astype(np.complex128)
    psi1 = evecs[:, 1].reshape(-1, 1).astype(np.complex128)
    if guess is None:
        guess = partial(_transmon_guess, T=T)
    S = partial(shapes.flattop, t_start=0.0, t_stop=T, t_rise=0.05 * T,
                func='sinsq')
    H = [H0, [H1, guess]]
    return Workload(
        name='C2_transmon_xgate_N%d' % (2 * nstates + 1),
        Hs=[H, H],
        initial_states=[psi0, psi1], targets=[psi1, psi0],
        pulse_options={guess: dict(lambda_a=lambda_a, update_shape=S)},
        tlist=np.linspace(0, T, nt), chi='re',
    )


# --- C3 ---------------------------------------------------------------------

def two_qubit_gate(nt=2000, T=25.0, w1=1.1, w2=2.1, J=0.2, la=1.1, u0=0.3,
                   lambda_a=100.0):
    """C3: two-qubit Hamiltonian of notebook 07 cell 12 with the four
    Bell-basis initial states of ``gate_objectives(..., 'PE')``
    (objectives.py:1044-1047).  The perfect-entangler functional needs the
    external ``weylchamber`` package, so the targets here are the images of
    those Bell states under sqrt(iSWAP) (a perfect entangler) and the
    functional is ``chis_sm``."""
    Hq1 = 0.5 * w1 * np.diag([-1, 1])
    Hq2 = 0.5 * w2 * np.diag([-1, 1])
    I2 = np.identity(2)
    H0 = np.kron(Hq1, I2) + np.kron(I2, Hq2)
    H0 = (H0 + 2 * J * (np.kron(_SX, _SX) + np.kron(_SY, _SY))).astype(
        np.complex128)
    H1 = (np.kron(_SX, I2) + la * np.kron(I2, _SX)).astype(np.complex128)
    guess = lambda t, args: u0 * shapes.flattop(  # noqa: E731
        t, t_start=0, t_stop=T, t_rise=T / 20, t_fall=T / 20, func='sinsq')
    S = partial(shapes.flattop, t_start=0, t_stop=T, t_rise=T / 20,
                t_fall=T / 20, func='sinsq')
    b = [_ket(4, i) for i in range(4)]
    r = np.sqrt(2)
    bell = [(b[0] + b[3]) / r, (1j * b[1] + 1j * b[2]) / r,
            (b[1] - b[2]) / r, (1j * b[0] - 1j * b[3]) / r]
    s = 1 / np.sqrt(2)
    gate = np.array([[1, 0, 0, 0], [0, s, 1j * s, 0], [0, 1j * s, s, 0],
                     [0, 0, 0, 1]], dtype=np.complex128)
    H = [H0, [H1, guess]]
    return Workload(
        name='C3_two_qubit_gate',
        Hs=[H] * 4,
        initial_states=bell, targets=[gate @ p for p in bell],
        pulse_options={guess: dict(lambda_a=lambda_a, update_shape=S)},
        tlist=np.linspace(0, T, nt), chi='sm',
    )


# --- C4 ---------------------------------------------------------------------

def tls_ensemble(K=128, nt=1000, T=5.0, ampl0=0.2, lambda_a=5.0,
                 omega_lo=0.9, omega_hi=1.1):
    """C4 (north star): K detuned two-level systems sharing one control,
    |0> -> |1>, ``chis_re``."""
    guess = lambda t, args: ampl0 * shapes.flattop(  # noqa: E731
        t, t_start=0, t_stop=T, t_rise=0.3, func='blackman')
    S = partial(shapes.flattop, t_start=0, t_stop=T, t_rise=0.3, t_fall=0.3,
                func='blackman')
    H1 = _SX.copy()
    omegas = np.linspace(omega_lo, omega_hi, K) if K > 1 else np.array([1.0])
    Hs = [[-0.5 * w * _SZ, [H1, guess]] for w in omegas]
    return Workload(
        name='C4_tls_ensemble_K%d' % K,
        Hs=Hs,
        initial_states=[_ket(2, 0)] * K, targets=[_ket(2, 1)] * K,
        pulse_options={guess: dict(lambda_a=lambda_a, update_shape=S)},
        tlist=np.linspace(0, T, nt), chi='re',
    )


# --- C5 ---------------------------------------------------------------------

def dissipative_qubit_reset(nt=5000, T=25.0, omega_q=1.0, omega_T=3.0,
                            J=0.1, kappa=0.04, beta=1.0, lambda_a=0.1):
    """C5: qubit + lossy TLS in Liouville space (super-operator 16x16),
    thermal state -> |00><00| with the qubit-only co-state of notebook 04
    cell 37 (``chi='qubit_reset'``)."""
    from .objectives import liouvillian
    I2 = np.identity(2)
    H0 = np.kron(0.5 * omega_q * np.diag([-1, 1]), I2) + np.kron(
        I2, 0.5 * omega_T * np.diag([-1, 1]))
    H0 = H0 + J * np.fliplr(np.diag([0, 1, 1, 0]))
    H1 = np.kron(0.5 * np.diag([-1, 1]), I2)
    Nth = 1.0 / (np.exp(beta * omega_T) - 1.0)
    L1 = np.sqrt(kappa * (Nth + 1)) * np.kron(I2, np.array([[0, 1], [0, 0]]))
    L2 = np.sqrt(kappa * Nth) * np.kron(I2, np.array([[0, 0], [1, 0]]))
    S = partial(shapes.flattop, t_start=0, t_stop=T, t_rise=0.05 * T,
                t_fall=0.05 * T, func='sinsq')
    guess = lambda t, args: (omega_T - omega_q) * S(t)  # noqa: E731
    L = liouvillian([H0.astype(np.complex128),
                     [H1.astype(np.complex128), guess]], [L1, L2])
    x_q, x_T = omega_q * beta / 2.0, omega_T * beta / 2.0
    rho_q = np.diag([np.exp(x_q), np.exp(-x_q)]) / (2 * np.cosh(x_q))
    rho_T = np.diag([np.exp(x_T), np.exp(-x_T)]) / (2 * np.cosh(x_T))
    rho_th = np.kron(rho_q, rho_T).astype(np.complex128)
    rho_trg = np.kron(np.diag([1, 0]), np.diag([1, 0])).astype(np.complex128)
    return Workload(
        name='C5_dissipative_qubit_reset',
        Hs=[L], initial_states=[rho_th], targets=[rho_trg],
        pulse_options={guess: dict(lambda_a=lambda_a, update_shape=S)},
        tlist=np.linspace(0, T, nt), chi='qubit_reset', is_super=True,
        meta=dict(chi_fixed=np.kron(np.diag([1, 0]), np.diag([1, 1])).astype(
            np.complex128)),
    )


# --- several controls (notebooks 02/03/08, tests/test_mu.py) ---------------

def _blackman_guess(t, args, ampl, t0, t1):
    return ampl * shapes.blackman(t, t_start=t0, t_stop=t1)


def _zero_guess(t, args):
    return 0.0


def lambda_system(nt=500, T=5.0, gamma=0.5, ensemble_mu=None, lambda_a=2.0,
                  E1=0.0, E2=10.0, E3=5.0, omega_P=9.5, omega_S=4.5,
                  ampl=5.0):
    """Λ-system in the rotating-wave approximation with FOUR real controls
    (real/imaginary parts of pump and Stokes pulse; N=3, L=4, M=5):
    docs/notebooks/03_example_lambda_system_rwa_non_hermitian.ipynb cells
    5-26 (``gamma`` > 0: decay -i gamma on the intermediate level, a
    non-Hermitian H0), notebook 02 (``gamma=0``), and with `ensemble_mu` the
    robustness ensemble of docs/notebooks/08_example_ensemble.ipynb cells
    20-39 (control Hamiltonians scaled by mu, all copies share the four
    control objects; notebook values: mu = 0.9 ... 1.1, lambda_a = 0.5).
    |1> -> e^{i (E2 - omega_S) T} |3>, ``chis_re``."""
    dP = E1 + omega_P - E2
    dS = E3 + omega_S - E2
    H0 = np.diag([dP, -1j * gamma, dS]).astype(np.complex128)
    HP_re = -0.5 * np.array([[0, 1, 0], [1, 0, 0], [0, 0, 0]], dtype=complex)
    HP_im = -0.5 * np.array([[0, 1j, 0], [-1j, 0, 0], [0, 0, 0]],
                            dtype=complex)
    HS_re = -0.5 * np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0]], dtype=complex)
    HS_im = -0.5 * np.array([[0, 0, 0], [0, 0, 1j], [0, -1j, 0]],
                            dtype=complex)
    P1 = partial(_blackman_guess, ampl=ampl, t0=2.0, t1=5.0)
    S1 = partial(_blackman_guess, ampl=ampl, t0=0.0, t1=3.0)
    # distinct control objects for the two zero guesses (controls are
    # identified by object identity, conversions.py:43-58)
    P2 = partial(_zero_guess)
    S2 = partial(_zero_guess)
    ctrl_ops = [(HP_re, P1), (HP_im, P2), (HS_re, S1), (HS_im, S2)]
    mus = [1.0] if ensemble_mu is None else list(ensemble_mu)
    Hs = [[H0] + [[mu * op, c] for op, c in ctrl_ops] for mu in mus]
    shape = partial(shapes.flattop, t_start=0.0, t_stop=T, t_rise=0.3,
                    func='sinsq')
    psi0 = _ket(3, 0)
    tgt = np.exp(1j * (E2 - omega_S) * T) * _ket(3, 2)
    return Workload(
        name='lambda_system_K%d' % len(mus),
        Hs=Hs, initial_states=[psi0] * len(mus), targets=[tgt] * len(mus),
        pulse_options={c: dict(lambda_a=lambda_a, update_shape=shape)
                       for _, c in ctrl_ops},
        tlist=np.linspace(0, T, nt), chi='re',
    )


def tls_shared_controls(nt=100, T=2.0):
    """The control system of the reference's tests/test_mu.py:8-27 as an
    optimisation problem: objective 0 contains control eps1 TWICE (sigma_+
    and sigma_- terms, so mu = sigma_x/2 + ... is a sum, mu.py:124) and not
    eps2; objective 1 contains only eps2 (a sigma_z term): the mapping of
    eps2 for objective 0 and of eps1 for objective 1 is empty (zero mu,
    mu.py:133-138).  N=2, L=2."""
    sp = np.array([[0, 1], [0, 0]], dtype=np.complex128)   # qutip sigmap()
    sm = sp.T.copy()
    eps1 = lambda t, args: 0.5   # noqa: E731
    eps2 = lambda t, args: 1.0   # noqa: E731
    H1 = [0.5 * _SZ, [sp, eps1], [sm, eps1]]
    H2 = [0.5 * _SZ, [_SZ.copy(), eps2]]
    shape = partial(shapes.flattop, t_start=0.0, t_stop=T, t_rise=0.2 * T,
                    func='sinsq')
    r = 1 / np.sqrt(2)
    return Workload(
        name='tls_shared_controls',
        Hs=[H1, H2],
        initial_states=[_ket(2, 0), r * (_ket(2, 0) + _ket(2, 1))],
        targets=[_ket(2, 1), r * (_ket(2, 0) - 1j * _ket(2, 1))],
        pulse_options={eps1: dict(lambda_a=1.0, update_shape=shape),
                       eps2: dict(lambda_a=2.0, update_shape=shape)},
        tlist=np.linspace(0, T, nt), chi='re',
    )


# --- large Liouville space (notebook 06) ------------------------------------

def two_transmon_gate(n_qubit=5, nt=2000, T=400.0, lambda_a=1.0):
    """Two coupled transmons with decay and dephasing in Liouville space, after
    docs/notebooks/06_example_3states.ipynb of the reference (cells 5-32):
    density matrices (n_qubit^2) x (n_qubit^2) -- super-operators 625 x 625
    for the notebook's n_qubit = 5 -- two real controls (the second one starts
    at zero), the three states of ``liouville_states_set='3states'`` as
    objectives, sqrt(iSWAP) as the target gate, ``chis_re``.  Units: GHz / ns
    with 2 pi folded into the frequencies."""
    from .objectives import gate_objectives, liouvillian
    GHz, MHz, ns = 2 * np.pi, 2 * np.pi * 1e-3, 1.0
    w1, w2, wd = 4.3796 * GHz, 4.6137 * GHz, 4.4985 * GHz
    d1, d2, J = -239.3 * MHz, -242.8 * MHz, -2.3 * MHz
    q1T1, q2T1, q1T2, q2T2 = 38.0e3 * ns, 32.0e3 * ns, 29.5e3 * ns, 16.0e3 * ns
    n = n_qubit
    b = np.diag(np.sqrt(np.arange(1, n)), 1).astype(np.complex128)
    I = np.identity(n)
    b1, b2 = np.kron(I, b), np.kron(b, I)
    n1, n2 = b1.conj().T @ b1, b2.conj().T @ b2
    H0 = ((w1 - wd - d1 / 2) * n1 + (d1 / 2) * n1 @ n1
          + (w2 - wd - d2 / 2) * n2 + (d2 / 2) * n2 @ n2
          + J * (b1.conj().T @ b2 + b1 @ b2.conj().T))
    H1_re = 0.5 * (b1 + b1.conj().T + b2 + b2.conj().T)
    H1_im = 0.5j * (b1.conj().T - b1 + b2.conj().T - b2)
    S = partial(shapes.flattop, t_start=0, t_stop=T, t_rise=20 * ns,
                t_fall=20 * ns, func='sinsq')
    omega = lambda t, args: 35.0 * MHz * S(t)  # noqa: E731
    zero = lambda t, args: 0.0  # noqa: E731
    c_ops = [np.sqrt(1 / q1T1) * b1, np.sqrt(1 / q2T1) * b2,
             np.sqrt(1 / q1T2) * n1, np.sqrt(1 / q2T2) * n2]
    L = liouvillian([H0, [H1_re, omega], [H1_im, zero]], c_ops)

    def ket(i, j):
        v = np.zeros((n * n, 1), dtype=np.complex128)
        v[i * n + j] = 1.0
        return v
    basis = [ket(0, 0), ket(0, 1), ket(1, 0), ket(1, 1)]
    sqrt_iswap = np.array([[1, 0, 0, 0],
                           [0, 1 / np.sqrt(2), 1j / np.sqrt(2), 0],
                           [0, 1j / np.sqrt(2), 1 / np.sqrt(2), 0],
                           [0, 0, 0, 1]], dtype=np.complex128)
    weights = np.array([20, 1, 1], dtype=np.float64)
    weights *= len(weights) / np.sum(weights)
    objs = gate_objectives(basis, sqrt_iswap, L, liouville_states_set='3states',
                           weights=weights, normalize_weights=False)
    return Workload(
        name='two_transmon_gate_N%d' % (n * n) ** 2,
        Hs=[o.H for o in objs], initial_states=[o.initial_state for o in objs],
        targets=[o.target for o in objs],
        pulse_options={omega: dict(lambda_a=lambda_a, update_shape=S),
                       zero: dict(lambda_a=lambda_a, update_shape=S)},
        tlist=np.linspace(0, T, nt), chi='re', is_super=True,
        weights=[float(o.weight) for o in objs],
    )


def by_name(name, **kwargs):
    """Workload from its BASELINE label ('C1'..'C5')."""
    table = {
        'C1': tls_state_to_state, 'C2': transmon_xgate, 'C3': two_qubit_gate,
        'C4': tls_ensemble, 'C5': dissipative_qubit_reset,
    }
    return table[name](**kwargs)
