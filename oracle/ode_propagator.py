"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's
``DensityMatrixODEPropagator`` (/root/reference/src/krotov/propagators.py:162-327)
and of the Krotov loop driven by it, for the PHYSICS-LEVEL check of SURVEY.md
§8 row a7.

The reference integrates d/dt vec(rho) = sum_n c_n L_n vec(rho) over every time
interval with SciPy's ``zvode`` (Adams, order <= 12, adaptive) and keeps ONE
integrator alive over a whole sweep: between two ``integrate`` calls only the
control coefficients in the argument list are swapped (propagators.py:249-253),
so the multistep history is carried across the switch.  The result therefore
depends on the integrator's internal step sequence and is NOT the exact
piecewise-constant propagation expm(L dt) (the
reference's own notebook 04 marks the output NBVAL_IGNORE_OUTPUT): the two
differ by ~2e-3 relative in the pulse after five iterations of notebook 04's
problem while reproducing the same printed digits.  The CUDA engine lowers the
propagator to the exact propagation, so parity for it is defined at that level:
tests compare the engine with THIS restatement on the digits notebook 04
prints (qubit error, int g_a, pulse range, tau) and bound the pulse deviation.

QuTiP's ``spmvpy_csr`` (un-vendored, qutip 4.7.x ``cy/spmatfuncs``) is restated
as ``out += coeff * (csr @ rho)``; ``mat2vec`` / ``vec2mat`` as column stacking.
Only ``tests/`` may import this module."""
import numpy as np
import scipy.integrate
import scipy.sparse

__all__ = ['DensityMatrixODEPropagator', 'optimize_single_control']


class DensityMatrixODEPropagator:
    """propagators.py:162-327 on column-stacked vectors: ``H`` is a list of
    ``[matrix, coefficient]`` pairs (coefficient 1 for drift terms, the pulse
    value otherwise); ``state`` a length-d^2 vector."""

    def __init__(self, method='adams', order=12, atol=1e-8, rtol=1e-6,
                 nsteps=1000, first_step=0, min_step=0, max_step=0):
        self.opts = dict(method=method, order=order, atol=atol, rtol=rtol,
                         nsteps=nsteps, first_step=first_step,
                         min_step=min_step, max_step=max_step)
        self._L_list = None
        self._r = None
        self._t = 0.0

    @staticmethod
    def _rhs(t, rho, L_list):
        # propagators.py:262-274 (spmvpy_csr accumulates coeff * L @ rho)
        out = np.zeros(rho.shape[0], dtype=complex)
        for L, coeff in L_list:
            out += coeff * (L @ rho)
        return out

    def __call__(self, H, state, dt, initialize=False):
        if initialize:
            # _initialize_data + _initialize_integrator (:276-327)
            self._L_list = [[scipy.sparse.csr_matrix(op), c] for op, c in H]
            r = scipy.integrate.ode(self._rhs)
            r.set_integrator('zvode', **self.opts)
            r.set_initial_value(np.asarray(state, dtype=complex))
            r.set_f_params(self._L_list)
            self._r = r
            self._t = 0.0
        else:
            # only the control values are swapped (:249-253)
            for i, (_, c) in enumerate(H):
                self._L_list[i][1] = c
        self._t += dt
        self._r.integrate(self._t)
        return np.array(self._r.y)


def optimize_single_control(L0, L1, rho0, chi_T, guess_pulse, shape, lambda_a,
                            tlist, iter_stop, make_propagator):
    """The loop of optimize.py:295-322, 393-508 for ONE objective in Liouville
    space with one control and a state-independent co-state (notebook 04's
    ``chis_qubit``), driven through a stateful propagator instance the way the
    reference drives it: ``initialize=True`` at the first step of every sweep
    (optimize.py:836, 882, 910), backward sweep under the adjoint generator
    with the conjugated (real) pulse (:868-870), mu = i L1 (mu.py:130-132),
    overlap tr(a^dag b).  Vectors are column-stacked density matrices.

    Returns a list of per-iteration dicts (index 0 = the guess): pulse
    (on the intervals), g_a, fw_state_T."""
    NT = len(tlist) - 1
    prop = make_propagator()
    L0a, L1a = L0.conj().T, L1.conj().T
    mu = 1j * L1
    pulse = np.array(guess_pulse, dtype=float)

    def forward(p):
        state = rho0
        for n in range(NT):
            dt = tlist[n + 1] - tlist[n]
            state = prop([[L0, 1], [L1, p[n]]], state, dt, initialize=(n == 0))
        return state

    records = [dict(pulse=pulse.copy(), g_a=0.0, fw_state_T=forward(pulse))]
    chi_norm = np.linalg.norm(chi_T)
    chi_unit = chi_T / chi_norm
    for _ in range(iter_stop):
        # backward sweep, all states stored (optimize.py:849-886)
        X = np.zeros((NT + 1, len(rho0)), dtype=complex)
        X[NT] = chi_unit
        state = chi_unit
        for n in range(NT - 1, -1, -1):
            dt = tlist[n + 1] - tlist[n]
            state = prop([[L0a, 1], [L1a, np.conjugate(pulse[n])]], state, dt,
                         initialize=(n == NT - 1))
            X[n] = state
        # update + forward sweep (optimize.py:449-500)
        new = pulse.copy()
        g_a = 0.0
        state = rho0
        for n in range(NT):
            dt = tlist[n + 1] - tlist[n]
            d1 = (chi_norm * np.vdot(X[n], mu @ state)).imag
            new[n] += (shape[n] / lambda_a) * d1
            g_a += (shape[n] / lambda_a) * abs(d1) ** 2 * dt
            state = prop([[L0, 1], [L1, new[n]]], state, dt,
                         initialize=(n == 0))
        pulse = new
        records.append(dict(pulse=pulse.copy(), g_a=g_a, fw_state_T=state))
    return records
