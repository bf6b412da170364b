#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the
UNMODIFIED reference package (/root/reference/src/krotov) in this container.

QuTiP, glom and grapheme are not installable here, so the reference is
imported through the stand-ins in oracle/ref_shims (a dense numpy-backed
``qutip.Qobj``; see its docstring) plus the alias
``np.ComplexWarning = np.exceptions.ComplexWarning`` that NumPy 2 requires for
/root/reference/src/krotov/conversions.py:103.  The reference's own
``optimize_pulses`` loop, ``propagators.expm``, ``mu.derivative_wrt_pulse``,
``second_order._overlap`` and ``functionals.chis_*`` then run as written
("qobj" path).  The "numpy" path feeds plain arrays with the user plugins of
docs/notebooks/09_example_numpy.ipynb cells 16/30/32.

/root/reference does not exist on the GPU box, so only the generated fixtures
travel.  Usage:  python oracle/make_golden.py [name ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, 'ref_shims'), '/root/reference/src', ROOT]
np.ComplexWarning = np.exceptions.ComplexWarning

import krotov  # noqa: E402  (the reference)
import qutip  # noqa: E402  (the shim)
import scipy.linalg  # noqa: E402

from krotov_b200 import workloads  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


# --- numpy-objects plugins, notebook 09 cells 16/30/32 ----------------------

def np_expm(H, state, dt, c_ops=None, backwards=False, initialize=False):
    eqm_factor = -1j
    if backwards:
        eqm_factor = eqm_factor.conjugate()
    A = eqm_factor * H[0]
    for part in H[1:]:
        A = A + (eqm_factor * part[1]) * part[0]
    return scipy.linalg.expm(A * dt) @ state


def np_overlap(a, b):
    if a is None or b is None:
        return None
    return complex(a.conj().T @ b)


def make_np_mu(objectives):
    def mu(objectives_, i_objective, pulses, pulses_mapping, i_pulse,
           time_index):
        idx = pulses_mapping[i_objective][0][i_pulse]

        def _mu(state):
            out = 0 * state
            for i in idx:
                out = out + objectives_[i_objective].H[i][0] @ state
            return out
        return _mu
    return mu


class Recorder:
    """info_hook capturing what the parity tests compare."""

    def __init__(self, keep_states):
        self.keep_states = keep_states
        self.pulses, self.g_a, self.tau, self.fwT = [], [], [], []
        self.bw = None
        self.fw = None

    @staticmethod
    def _arr(x):
        return x.full() if hasattr(x, 'full') else np.asarray(x)

    def __call__(self, **kw):
        self.pulses.append(np.array([p.copy() for p in kw['optimized_pulses']]))
        self.g_a.append(np.array(kw['g_a_integrals']).copy())
        tau = kw['tau_vals']
        self.tau.append(np.array([complex(t) for t in tau]))
        self.fwT.append(np.array(
            [self._arr(s).reshape(-1, order='F') for s in kw['fw_states_T']]))
        if self.keep_states and kw['iteration'] == 1:
            self.bw = np.array([
                [self._arr(s).reshape(-1, order='F') for s in states]
                for states in kw['backward_states']])
            if kw['forward_states'] is not None:
                self.fw = np.array([
                    [self._arr(s).reshape(-1, order='F') for s in states]
                    for states in kw['forward_states']])
        return None


CHI = {
    're': krotov.functionals.chis_re,
    'ss': krotov.functionals.chis_ss,
    'sm': krotov.functionals.chis_sm,
    'hs': krotov.functionals.chis_hs,
}


def run_reference(wl, iters, path, keep_states, sigma=None, modify=None):
    krotov.Objective.type_checking = (path == 'qobj')
    wrap = (lambda a: qutip.Qobj(a)) if path == 'qobj' else None
    if wl.is_super and path == 'qobj':
        d = wl.initial_states[0].shape[0]

        def wrap(a):  # noqa: F811
            if a.shape == (d * d, d * d) and d > 1:
                return qutip.Qobj(a, dims=[[[d], [d]], [[d], [d]]])
            return qutip.Qobj(a)
    objectives = wl.objectives(krotov.Objective, wrap=wrap)
    if wl.chi == 'qubit_reset':
        fixed = wl.meta['chi_fixed']

        def chi_constructor(fw_states_T, objectives, tau_vals):
            return [wrap(fixed) if wrap else fixed.copy()
                    for _ in fw_states_T]
    else:
        chi_constructor = CHI[wl.chi]
    rec = Recorder(keep_states)
    kwargs = {}
    if path == 'numpy':
        kwargs = dict(propagator=np_expm, mu=make_np_mu(objectives),
                      overlap=np_overlap, norm=np.linalg.norm)
    else:
        kwargs = dict(propagator=krotov.propagators.expm)
    t0 = time.time()
    res = krotov.optimize_pulses(
        objectives, pulse_options=wl.pulse_options, tlist=wl.tlist,
        chi_constructor=chi_constructor, info_hook=rec, iter_stop=iters,
        sigma=sigma, modify_params_after_iter=modify, **kwargs)
    secs = time.time() - t0
    out = dict(
        tlist=wl.tlist,
        guess_pulses=rec.pulses[0],
        pulses=np.array(rec.pulses),      # [iters+1][L][NT]
        g_a=np.array(rec.g_a),            # [iters+1][L]
        tau=np.array(rec.tau),            # [iters+1][K]
        fw_states_T=np.array(rec.fwT),    # [iters+1][K][N]
        optimized_controls=np.array(res.optimized_controls),
        seconds=np.array(secs),
    )
    if rec.bw is not None:
        out['backward_states_it1'] = rec.bw
    if rec.fw is not None:
        out['forward_states_it1'] = rec.fw
    return out


def save(name, **arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + '.npz')
    np.savez_compressed(path, **arrays)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


# --- cases -------------------------------------------------------------------

def case_tls_fixture():
    """tests/test_krotov.py fixture; rows of tests/test_krotov/oct.log."""
    wl = workloads.tls_reference_fixture()
    out = run_reference(wl, 3, 'qobj', keep_states=True)
    out_np = run_reference(wl, 3, 'numpy', keep_states=False)
    out['pulses_numpy_path'] = out_np['pulses']
    save('tls_fixture_qobj', **out)


def case_c1():
    wl = workloads.tls_state_to_state()
    save('C1_qobj', **run_reference(wl, 3, 'qobj', keep_states=True))


def case_c2():
    wl = workloads.transmon_xgate(nstates=1, nt=1000)
    save('C2_qobj', **run_reference(wl, 3, 'qobj', keep_states=False))
    wl = workloads.transmon_xgate(nstates=2, nt=100)
    save('transmon_N5_nt100_qobj',
         **run_reference(wl, 2, 'qobj', keep_states=True))
    wl = workloads.transmon_xgate(nstates=8, nt=200)
    save('transmon_N17_nt200_qobj',
         **run_reference(wl, 2, 'qobj', keep_states=False))


class ConstSigma(krotov.second_order.Sigma):
    """sigma(t) = -max(0, 2A) with A re-estimated by the reference's own
    ``numerical_estimate_A`` (docs/notebooks/07_example_PE.ipynb cell 30),
    Delta J_T from J_T_sm."""

    def __init__(self, A):
        self.A = A
        self.A_hist = [A]
        self._J_prev = None

    def __call__(self, t):
        return -max(0.0, 2 * self.A)

    def refresh(self, forward_states, forward_states0, chi_states, chi_norms,
                optimized_pulses, guess_pulses, objectives, result):
        taus = result.tau_vals
        J1 = 1 - abs(np.mean(taus[-1])) ** 2
        J0 = 1 - abs(np.mean(taus[-2])) ** 2
        self.A = krotov.second_order.numerical_estimate_A(
            forward_states, forward_states0, chi_states, chi_norms, J1 - J0)
        self.A_hist.append(self.A)


def case_c3():
    wl = workloads.two_qubit_gate(nt=250)
    save('C3_nt250_first_order_qobj',
         **run_reference(wl, 3, 'qobj', keep_states=False))
    sig = ConstSigma(A=0.5)
    out = run_reference(wl, 3, 'qobj', keep_states=True, sigma=sig)
    out['sigma_A'] = np.array(sig.A_hist)
    save('C3_nt250_second_order_qobj', **out)


def case_c4():
    wl = workloads.tls_ensemble(K=8, nt=200)
    save('C4_K8_nt200_qobj', **run_reference(wl, 3, 'qobj', keep_states=True))
    wl = workloads.tls_ensemble(K=128, nt=1000)
    out = run_reference(wl, 2, 'numpy', keep_states=False)
    save('C4_K128_nt1000_numpy', **out)


def case_c5():
    wl = workloads.dissipative_qubit_reset(nt=500)
    save('C5_nt500_qobj', **run_reference(wl, 3, 'qobj', keep_states=True))


def case_multi_control():
    """Several controls per objective (L > 1, M > 2): the non-Hermitian
    Lambda system of notebook 03, the 5-member Lambda ensemble of notebook 08
    and the repeated / absent control system of tests/test_mu.py:8-27."""
    wl = workloads.lambda_system(nt=500, gamma=0.5)
    save('lambda_nonherm_qobj', **run_reference(wl, 3, 'qobj',
                                                 keep_states=True))
    wl = workloads.lambda_system(nt=500, gamma=0.0, lambda_a=0.5,
                                 ensemble_mu=[0.9, 0.95, 1.0, 1.05, 1.1])
    save('lambda_ensemble_qobj', **run_reference(wl, 3, 'qobj',
                                                  keep_states=False))
    wl = workloads.tls_shared_controls()
    save('shared_controls_qobj', **run_reference(wl, 3, 'qobj',
                                                  keep_states=True))


def case_full_size():
    """C3 and C5 at the sizes BASELINE.json's configs name (nt=2000 / 5000);
    pulses, tau and final states only."""
    wl = workloads.two_qubit_gate(nt=2000)
    save('C3_nt2000_first_order_qobj',
         **run_reference(wl, 3, 'qobj', keep_states=False))
    sig = ConstSigma(A=0.5)
    out = run_reference(wl, 3, 'qobj', keep_states=False, sigma=sig)
    out['sigma_A'] = np.array(sig.A_hist)
    save('C3_nt2000_second_order_qobj', **out)
    wl = workloads.dissipative_qubit_reset(nt=5000)
    save('C5_nt5000_qobj', **run_reference(wl, 2, 'qobj', keep_states=False))


def case_infohook_kat():
    """tests/test_infohooks.py:15-72: lambda halved after each iteration;
    golden info_vals[1][0] = 0.001978333994757067."""
    import scipy
    Ec, EjEc, nstates, ng, T = 0.386, 45, 2, 0.0, 10.0
    Ej = EjEc * Ec
    n = np.arange(-nstates, nstates + 1)
    up = np.diag(np.ones(2 * nstates), k=-1)
    do = up.T
    H0 = qutip.Qobj(np.diag(4 * Ec * (n - ng) ** 2) - Ej * (up + do) / 2.0)
    H1 = qutip.Qobj(-2 * np.diag(n))
    eigenvals, eigenvecs = scipy.linalg.eig(H0.full())
    ndx = np.argsort(eigenvals.real)
    E = eigenvals[ndx].real
    V = eigenvecs[:, ndx]
    w01 = E[1] - E[0]
    psi0 = qutip.Qobj(V[:, 0])
    psi1 = qutip.Qobj(V[:, 1])
    profile = lambda t: np.exp(-40.0 * (t / T - 0.5) ** 2)  # noqa: E731
    eps0 = lambda t, args: 0.5 * profile(t) * np.cos(  # noqa: E731
        8 * np.pi * w01 * t)
    H = [H0, [H1, eps0]]
    krotov.Objective.type_checking = True
    obj = krotov.Objective(initial_state=psi0, target=psi1, H=H)
    tlist = np.array([0, 0.01, 0.02])

    def adjust(**args):
        args['lambda_vals'][0] *= 0.5

    rec = Recorder(False)

    def fid(**args):
        rec(**args)
        return np.average(np.array(args['tau_vals']).real)

    res = krotov.optimize_pulses(
        [obj], pulse_options={H[1][1]: dict(lambda_a=1, update_shape=1)},
        tlist=tlist, propagator=krotov.propagators.expm,
        chi_constructor=krotov.functionals.chis_re, info_hook=fid,
        modify_params_after_iter=adjust, iter_stop=2)
    assert abs(res.info_vals[1] - 0.001978333994757067) < 1e-8
    save('infohook_kat_qobj', H0=H0.full(), H1=H1.full(), psi0=psi0.full(),
         psi1=psi1.full(), tlist=tlist, w01=np.array(w01),
         guess_pulses=rec.pulses[0], pulses=np.array(rec.pulses),
         tau=np.array(rec.tau), info_vals=np.array(res.info_vals),
         g_a=np.array(rec.g_a))


def case_large_liouville():
    """Row f3: the two-transmon Liouville problem of notebook 06 (three
    weighted objectives, two controls, chis_re) with three levels per transmon
    (super-operators 81 x 81) and with the notebook's five (625 x 625) on a
    short grid."""
    wl = workloads.two_transmon_gate(n_qubit=3, nt=60, T=12.0)
    save('two_transmon_N81_qobj', **run_reference(wl, 2, 'qobj',
                                                   keep_states=True))
    wl = workloads.two_transmon_gate(n_qubit=5, nt=9, T=1.8)
    save('two_transmon_N625_qobj', **run_reference(wl, 1, 'qobj',
                                                    keep_states=False))


CASES = dict(tls_fixture=case_tls_fixture, c1=case_c1, c2=case_c2,
             c3=case_c3, c4=case_c4, c5=case_c5, kat=case_infohook_kat,
             multi=case_multi_control, full=case_full_size,
             large=case_large_liouville)

if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        t0 = time.time()
        CASES[nm]()
        print("case %s done in %.1f s" % (nm, time.time() - t0))
