"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Krotov sweep hot path.

A plain numpy/scipy restatement of the reference's algorithm for the path
SURVEY.md §8 scopes (one Krotov iteration = chi boundary -> backward sweep ->
sequential pulse-update/forward sweep -> tau), written on dense arrays.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package
``krotov_b200`` never does (it fails loudly when its CUDA library is absent).

Each function cites the reference lines it follows (paths relative to
/root/reference).  The arithmetic that the reference delegates to un-vendored
third parties is restated the way the reference's own numpy recipe does
(docs/notebooks/09_example_numpy.ipynb cells 16/30/32):

* QuTiP 4.x (``qutip>=4.3.1,<5.0``, pyproject.toml:33; 4.7.6 produced all
  recorded outputs) ``Qobj.expm`` -> ``scipy.linalg.expm`` (Pade + scaling and
  squaring, Al-Mohy & Higham 2009) of the dense generator,
  ``Qobj.__call__`` -> matrix-vector product (column-stacking ``vec`` for
  super-operators acting on density matrices), ``Qobj.overlap`` /
  ``tr(a^dag b)`` -> ``vdot``, ``Qobj.norm`` -> L2 / Frobenius.
* SciPy 1.18.1 (installed in this image) provides ``scipy.linalg.expm``.

PARITY PINNING: tests/test_oracle.py checks this module against golden
vectors produced by running the *unmodified* reference package
(/root/reference/src/krotov) in this container through import shims
(oracle/ref_shims, script oracle/make_golden.py, fixtures tests/golden/*.npz)
-- both on the reference's Qobj code path (its own ``propagators.expm``,
``mu.derivative_wrt_pulse``, ``_overlap``, ``functionals.chis_*``) and on its
numpy-objects path -- and against the reference's own known-answer values
(tests/test_infohooks.py:67, tests/test_krotov/oct.log,
tests/test_parallelization.py:139-140).

States are 1-D complex128 vectors of length N (kets, or column-stacked
density matrices).  ``terms[k]`` is a list of ``(matrix[N,N], pulse_index)``
with ``pulse_index = -1`` for drift terms (coefficient 1).
"""
import numpy as np
import scipy.linalg

__all__ = [
    'control_onto_interval', 'pulse_onto_tlist', 'discretize_via_midpoints',
    'generator', 'expm_step', 'forward_propagation', 'backward_propagation',
    'mu_operator', 'krotov_iteration', 'optimize', 'chis_re', 'chis_ss',
    'chis_sm', 'chis_hs',
]


# --- control discretisation (conversions.py:61-137, 333-390) ---------------

def control_onto_interval(control):
    """conversions.py:333-364: sequential un-averaging recurrence."""
    control = np.asarray(control, dtype=np.float64)
    pulse = np.zeros(len(control) - 1)
    pulse[0] = control[0]
    for i in range(1, len(control) - 1):
        pulse[i] = 2.0 * control[i] - pulse[i - 1]
    pulse[-1] = control[-1]
    return pulse


def pulse_onto_tlist(pulse):
    """conversions.py:368-390."""
    pulse = np.asarray(pulse, dtype=np.float64)
    control = np.zeros(len(pulse) + 1)
    control[0] = pulse[0]
    for i in range(1, len(control) - 1):
        control[i] = 0.5 * (pulse[i - 1] + pulse[i])
    control[-1] = pulse[-1]
    return control


def discretize_via_midpoints(func, tlist, args=()):
    """conversions.py:107-119 (callable control, ``via_midpoints=True``)."""
    mid = (tlist + 0.5 * (tlist[1] - tlist[0]))[:-1]
    mid[0] = tlist[0]
    mid[-1] = tlist[-1]
    samples = np.array([float(func(t, *args)) for t in mid])
    return pulse_onto_tlist(samples)


# --- single-step propagation (propagators.py:79-122) -----------------------

def generator(terms_k, pulses, n, factor, conjugate=False):
    """``A = f*Op_0 + sum (f*c_m)*Op_m`` in the reference's order of
    operations (propagators.py:94-111); pulse values are plugged in as in
    conversions.py:288-330 (conjugated for the backward sweep)."""
    A = None
    for op, l in terms_k:
        if l < 0:
            part = factor * op
        else:
            c = pulses[l][n]
            if conjugate:
                c = np.conjugate(c)
            part = (factor * c) * op
        A = part if A is None else A + part
    return A


def expm_step(terms_k, pulses, n, dt, state, is_super, backwards):
    """One call of ``propagators.expm``: factor -i (Hilbert) or 1
    (super-operator), conjugated for ``backwards`` (propagators.py:94-105);
    result ``expm(A*dt) @ state`` (:116-117)."""
    factor = 1 if is_super else -1j
    if backwards:
        factor = np.conjugate(factor)
    A = generator(terms_k, pulses, n, factor, conjugate=backwards)
    return scipy.linalg.expm(A * dt) @ state


def forward_propagation(terms_k, pulses, tlist, psi0, is_super,
                        store_all=True):
    """optimize.py:806-846."""
    nt = len(tlist)
    state = np.asarray(psi0, dtype=np.complex128)
    out = np.zeros((nt, len(state)), dtype=np.complex128) if store_all else None
    if store_all:
        out[0] = state
    for n in range(nt - 1):
        dt = tlist[n + 1] - tlist[n]
        state = expm_step(terms_k, pulses, n, dt, state, is_super, False)
        if store_all:
            out[n + 1] = state
    return out if store_all else state


def backward_propagation(adj_terms_k, pulses, tlist, chi_T, is_super):
    """optimize.py:849-886: propagate under the adjoint generator with
    ``backwards=True``; ``storage[n]`` holds chi at ``tlist[n]``."""
    nt = len(tlist)
    state = np.asarray(chi_T, dtype=np.complex128)
    out = np.zeros((nt, len(state)), dtype=np.complex128)
    out[-1] = state
    for n in range(nt - 2, -1, -1):
        dt = tlist[n + 1] - tlist[n]
        state = expm_step(adj_terms_k, pulses, n, dt, state, is_super, True)
        out[n] = state
    return out


def adjoint_terms(terms_k):
    """objectives.py:51-93 / 240-258: element-wise adjoint of every operator,
    controls untouched."""
    return [(op.conj().T, l) for op, l in terms_k]


def mu_operator(terms_k, l, is_super):
    """mu.py:123-140: sum of the operators driven by pulse ``l`` (times i for
    super-operators); None when the pulse does not occur (zero map)."""
    ops = [op for op, ll in terms_k if ll == l]
    if not ops:
        return None
    if is_super:
        mu = ops[0] * 1j
        for op in ops[1:]:
            mu = mu + (1j * 1) * op
    else:
        mu = ops[0].copy()
        for op in ops[1:]:
            mu = mu + (1j * -1j) * op
    return mu


# --- chi constructors (functionals.py:177-197, 225-253, 293-317, 389-437) --

def _weights(weights, K):
    return [None] * K if weights is None else list(weights)


def chis_re(fw_states_T, targets, tau_vals, weights=None):
    K = len(targets)
    c = 1.0 / (2 * K)
    return [
        (c * t if w is None else c * w * t)
        for t, w in zip(targets, _weights(weights, K))
    ]


def chis_ss(fw_states_T, targets, tau_vals, weights=None):
    K = len(targets)
    return [
        ((tau / K) * t if w is None else (tau / K) * w * t)
        for t, tau, w in zip(targets, tau_vals, _weights(weights, K))
    ]


def chis_sm(fw_states_T, targets, tau_vals, weights=None):
    K = len(targets)
    s = 0
    for tau, w in zip(tau_vals, _weights(weights, K)):
        s += tau if w is None else w * tau
    c = 1.0 / K ** 2
    return [
        (c * t * s if w is None else c * w * t * s)
        for t, w in zip(targets, _weights(weights, K))
    ]


def chis_hs(fw_states_T, targets, tau_vals, weights=None):
    K = len(targets)
    c = 1.0 / (2 * K)
    return [
        (c * (t - phi) if w is None else c * w * (t - phi))
        for t, phi, w in zip(targets, fw_states_T, _weights(weights, K))
    ]


def state_norm(state, is_super, operator_norm='trace'):
    """Default ``norm`` of optimize.py:242-243, i.e. ``Qobj.norm()``: L2 for
    kets, trace norm for density matrices / operators.  ``operator_norm='fro'``
    selects the Frobenius norm instead (what the CUDA engine uses; the update
    is invariant because chi/||chi|| is propagated linearly and ||chi|| is
    multiplied back, optimize.py:410,467)."""
    if is_super and operator_norm == 'trace':
        d = int(round(np.sqrt(len(state))))
        return float(np.sum(scipy.linalg.svdvals(
            state.reshape(d, d, order='F'))))
    return float(np.linalg.norm(state))


# --- one Krotov iteration (optimize.py:393-508) ----------------------------

def krotov_iteration(terms, psi0, targets, guess_pulses, shapes, lambdas,
                     tlist, fw_states_T, tau_vals, chi_constructor, is_super,
                     sigma=None, forward_states0=None, weights=None,
                     operator_norm='trace'):
    """Restatement of the main-loop body, optimize.py:393-508.

    Returns a dict with ``optimized_pulses`` (list of L arrays), ``g_a``,
    ``tau_vals``, ``fw_states_T``, ``backward_states`` ([K][nt][N]),
    ``forward_states`` (second order only), ``chi_norms``, ``chi_states``.
    """
    K = len(terms)
    L = len(guess_pulses)
    nt = len(tlist)
    second_order = sigma is not None
    # boundary condition and normalisation, optimize.py:404-410
    chis = chi_constructor(fw_states_T, targets, tau_vals, weights)
    chi_norms = [state_norm(c, is_super, operator_norm) for c in chis]
    chis = [c / nrm for c, nrm in zip(chis, chi_norms)]
    # backward sweep under the guess pulses, optimize.py:413-425
    adj = [adjoint_terms(t) for t in terms]
    backward_states = [
        backward_propagation(adj[k], guess_pulses, tlist, chis[k], is_super)
        for k in range(K)
    ]
    mus = [[mu_operator(terms[k], l, is_super) for l in range(L)]
           for k in range(K)]
    # sequential update + forward sweep, optimize.py:429-500
    g_a = np.zeros(L)
    optimized = [p.copy() for p in guess_pulses]
    fw = [np.asarray(p, dtype=np.complex128) for p in psi0]
    forward_states = None
    if second_order:
        forward_states = [np.zeros((nt, len(fw[k])), dtype=np.complex128)
                          for k in range(K)]
        for k in range(K):
            forward_states[k][0] = fw[k]
        delta_phis = [np.zeros_like(fw[k]) for k in range(K)]
    for n in range(nt - 1):
        dt = tlist[n + 1] - tlist[n]
        if second_order:
            sig = sigma(tlist[n] + 0.5 * dt)
        for l in range(L):
            acc = 0j
            for k in range(K):
                if mus[k][l] is None:
                    mu_psi = 0 * fw[k]
                else:
                    mu_psi = mus[k][l] @ fw[k]
                update = complex(np.vdot(backward_states[k][n], mu_psi))
                update *= chi_norms[k]
                if second_order:
                    update += 0.5 * sig * complex(
                        np.vdot(delta_phis[k], mu_psi))
                acc += update
            S_t = shapes[l][n]
            d1 = acc.imag
            delta = (S_t / lambdas[l]) * d1
            g_a[l] += (S_t / lambdas[l]) * abs(d1) ** 2 * dt
            optimized[l][n] += delta
        fw = [
            expm_step(terms[k], optimized, n, dt, fw[k], is_super, False)
            for k in range(K)
        ]
        if second_order:
            delta_phis = [fw[k] - forward_states0[k][n + 1] for k in range(K)]
            for k in range(K):
                forward_states[k][n + 1] = fw[k]
    tau = None
    if targets is not None and all(
            isinstance(t, np.ndarray) for t in targets):
        tau = np.array([complex(np.vdot(targets[k], fw[k])) for k in range(K)])
    return dict(optimized_pulses=optimized, g_a=g_a, tau_vals=tau,
                fw_states_T=fw, backward_states=backward_states,
                forward_states=forward_states, chi_norms=chi_norms,
                chi_states=chis)


def optimize(terms, psi0, targets, guess_pulses, shapes, lambdas, tlist,
             chi_constructor, is_super=False, iter_stop=1, sigma=None,
             sigma_refresh=None, weights=None, lambda_schedule=None,
             operator_norm='trace'):
    """Restatement of ``optimize_pulses`` without bookkeeping
    (optimize.py:295-322 initial forward sweep, :393-577 loop).

    ``sigma_refresh(record, forward_states0)`` is called after every
    iteration like ``Sigma.refresh`` (:566-577).  ``lambda_schedule(i)`` may
    return new lambda values after iteration ``i`` (what
    ``modify_params_after_iter`` does in tests/test_infohooks.py:30-37).

    Returns a list of per-iteration records (index 0 = iteration 0).
    """
    K = len(terms)
    lambdas = np.array(lambdas, dtype=np.float64)
    pulses = [np.array(p, dtype=np.float64) for p in guess_pulses]
    fw_all = [forward_propagation(terms[k], pulses, tlist, psi0[k], is_super)
              for k in range(K)]
    fw_T = [f[-1] for f in fw_all]
    tau = np.array([complex(np.vdot(targets[k], fw_T[k])) for k in range(K)])
    records = [dict(optimized_pulses=[p.copy() for p in pulses],
                    g_a=np.zeros(len(pulses)), tau_vals=tau, fw_states_T=fw_T,
                    forward_states=fw_all)]
    forward_states0 = fw_all if sigma is not None else None
    if lambda_schedule is not None:
        new = lambda_schedule(0)
        if new is not None:
            lambdas = np.array(new, dtype=np.float64)
    for it in range(1, iter_stop + 1):
        rec = krotov_iteration(terms, psi0, targets, pulses, shapes, lambdas,
                               tlist, fw_T, tau, chi_constructor, is_super,
                               sigma=sigma, forward_states0=forward_states0,
                               weights=weights, operator_norm=operator_norm)
        records.append(rec)
        if it >= iter_stop:
            break  # optimize.py:552-556: no refresh after the last iteration
        if sigma is not None and sigma_refresh is not None:
            sigma_refresh(rec, forward_states0, pulses)
        pulses = [p.copy() for p in rec['optimized_pulses']]
        fw_T, tau = rec['fw_states_T'], rec['tau_vals']
        if sigma is not None:
            forward_states0 = rec['forward_states']
        if lambda_schedule is not None:
            new = lambda_schedule(it)
            if new is not None:
                lambdas = np.array(new, dtype=np.float64)
    return records
