"""Problem compiler: lowers the reference's nested-list problem description
to the dense arrays the sm_100a sweep kernels consume.

What the reference redoes at *every time step* in Python -- copy the nested
list, plug in the pulse values (conversions.py:288-330), sum the operators
(propagators.py:100-111), rebuild mu (mu.py:123-140), take adjoints
(objectives.py:240-258) -- is done here ONCE per ``optimize_pulses`` call:

* ``ops[k][m]``      generator terms of objective k (term 0 = sum of all
  drift operators, one further term per distinct pulse driving k),
* ``ops_adj[k][m]``  their adjoints (backward generator),
* ``mu[k][l]``       dH/d eps_l (times i for super-operators),
* ``term2pulse[k][m]``, ``op_norm[k][m]`` (1-norms for the Taylor plan),
* discretised guess pulses / update shapes on the time *intervals*
  (optimize.py:641-704), ``dt[n] = tlist[n+1]-tlist[n]`` (optimize.py:450).

Matrices are stored column-major (element (r,c) at ``c*N+r``), states as
vectors of length N (column-stacked density matrices for Liouville space),
see include/krotov_b200.h.
"""
import numpy as np

from . import shapes as _shapes
from ._dense import dense, kind_of
from .conversions import (control_onto_interval, discretize, extract_controls,
                          extract_controls_mapping,
                          pulse_options_dict_to_list)

__all__ = ['CompiledProblem', 'compile_problem', 'initialize_controls']


def _shape_callable(val):
    if callable(val):
        return val
    if val == 1:
        return _shapes.one_shape
    if val == 0:
        return _shapes.zero_shape
    raise ValueError("update_shape must be a callable")


def _clip_shape(arr):
    """Values in [0,1] with a ±0.01 rounding margin, then clipped
    (optimize.py:605-620)."""
    lo, hi = np.min(arr), np.max(arr)
    if lo < -0.01 or hi > 1.01:
        raise ValueError(
            "Update shapes ('update_shape' in pulse options-dict) must have "
            "values in the range [0, 1], not [%s, %s]" % (lo, hi)
        )
    return np.clip(arr, 0.0, 1.0)


def initialize_controls(objectives, pulse_options, tlist):
    """Controls → discretised guess controls/pulses, mapping, lambda values
    and shape arrays; same results and errors as
    ``_initialize_krotov_controls`` (optimize.py:641-704)."""
    controls = extract_controls(objectives)
    mapping = extract_controls_mapping(objectives, controls)
    options = pulse_options_dict_to_list(pulse_options, controls)
    try:
        guess_controls = [
            discretize(c, tlist, args=(options[i].get('args', None),),
                       via_midpoints=True)
            for i, c in enumerate(controls)
        ]
    except (TypeError, np.exceptions.ComplexWarning) as exc:
        raise ValueError(
            "Cannot discretize controls: %s. Note that "
            "all controls must be real-valued. Complex controls must be "
            "split into an independent real and imaginary part in the "
            "objectives before passing them to the optimization" % exc
        )
    guess_pulses = [control_onto_interval(c) for c in guess_controls]
    try:
        lambda_vals = np.array([float(o['lambda_a']) for o in options])
    except KeyError:
        raise ValueError(
            "Each value in pulse_options must be a dict that contains "
            "the key 'lambda_a'."
        )
    shape_arrays = []
    for o in options:
        try:
            S = discretize(_shape_callable(o['update_shape']), tlist,
                           args=(), via_midpoints=True)
        except KeyError:
            raise ValueError(
                "Each value in pulse_options must be a dict that contains "
                "the key 'update_shape'."
            )
        except (TypeError, np.exceptions.ComplexWarning) as exc:
            raise ValueError(
                "Update shapes ('update_shape' in pulse options-dict) must be "
                "real-valued: %s" % exc
            )
        shape_arrays.append(_clip_shape(control_onto_interval(S)))
    return (controls, guess_controls, guess_pulses, mapping, lambda_vals,
            shape_arrays)


class CompiledProblem:
    """Dense, kernel-ready form of a list of objectives (host arrays)."""

    def __init__(self):
        self.K = self.N = self.NT = self.L = self.M = 0
        self.is_super = False
        self.real_ops = False
        self.ops = self.ops_adj = self.mu = None      # complex128 arrays
        self.sparse = None      # CSR bundle instead, for N > DENSE_NMAX
        self.term2pulse = self.op_norm = None
        self.row_nnz = 0
        self.psi0 = self.targets = None               # [K,N]
        self.weights = None
        self.dt = None
        self.state_templates = None
        self.state_shape = None

    def vec(self, state):
        """Quantum object → vector of length N (column-stacking)."""
        a = dense(state)
        if self.is_super:
            return a.reshape(-1, order='F')
        return a.reshape(-1)

    def unvec(self, v, template):
        """Vector → object shaped like `template` (inverse of :meth:`vec`)."""
        from ._dense import like
        v = np.asarray(v)
        if self.is_super:
            d = int(round(np.sqrt(self.N)))
            return like(template, v.reshape(d, d, order='F'))
        return like(template, v.reshape(self.state_shape))


def _classify(objective):
    """(N, is_super, state_shape) of one objective; raises like
    propagators.expm (propagators.py:112-122) for unsupported combinations."""
    H = objective.H if isinstance(objective.H, list) else [objective.H]
    op0 = H[0][0] if isinstance(H[0], list) else H[0]
    a = dense(op0)
    s = dense(objective.initial_state)
    skind = kind_of(objective.initial_state)
    okind = kind_of(op0)
    if a.shape[0] != a.shape[1]:
        raise ValueError("generator must be a square matrix")
    if skind in ('ket', 'bra') or (s.shape[1] == 1 and okind != 'super'):
        if a.shape[0] != s.size:
            raise ValueError(
                "dimension mismatch: H is %dx%d, state has %d elements"
                % (a.shape[0], a.shape[1], s.size))
        return s.size, False, s.shape
    # operator-valued state (density matrix)
    d = s.shape[0]
    if s.shape[0] != s.shape[1]:
        raise ValueError("state must be a ket or a square density matrix")
    if okind == 'super' or a.shape[0] == d * d and d > 1:
        if a.shape[0] != d * d:
            raise ValueError("super-operator/state dimension mismatch")
        return d * d, True, s.shape
    raise NotImplementedError(
        "Cannot handle argument types A:%s, state:%s" % ('oper', 'oper'))


def compile_problem(objectives, controls, mapping, tlist, mu=None,
                    pulses_for_mu=None):
    """Lower `objectives` (see module docstring).

    Args:
        objectives: list of :class:`krotov_b200.Objective`.
        controls, mapping: from :func:`initialize_controls`.
        tlist: time grid.
        mu: None / ``derivative_wrt_pulse`` for the standard linear-control
            derivative, or a custom callable with the reference's ``mu``
            signature; a custom ``mu`` is evaluated once per (objective,
            pulse) -- it must be linear in the state and independent of time
            and pulse values, otherwise NotImplementedError is raised.

    Raises:
        NotImplementedError: collapse operators (propagators.py:90-91),
            mixed Hilbert/Liouville objectives, unequal dimensions, or an
            un-lowerable ``mu``.
    """
    from .mu import derivative_wrt_pulse
    cp = CompiledProblem()
    K = len(objectives)
    if K == 0:
        raise ValueError("no objectives")
    L = len(controls)
    infos = []
    for obj in objectives:
        if len(obj.c_ops) > 0:
            raise NotImplementedError(
                "Liouville exponentiation not implemented")
        infos.append(_classify(obj))
    N, is_super, sshape = infos[0]
    for info in infos[1:]:
        if info[0] != N or info[1] != is_super:
            raise NotImplementedError(
                "all objectives must share the state dimension and the "
                "Hilbert/Liouville space type on the B200 engine")
    cp.K, cp.N, cp.L, cp.NT = K, N, L, len(tlist) - 1
    cp.is_super = is_super
    cp.state_shape = sshape
    tl = np.asarray(tlist, dtype=np.float64)
    cp.dt = np.array([tl[n + 1] - tl[n] for n in range(len(tl) - 1)])

    # generator terms: drift sum + one term per distinct pulse
    per_obj = []
    for k, obj in enumerate(objectives):
        H = obj.H if isinstance(obj.H, list) else [obj.H]
        drift = np.zeros((N, N), dtype=np.complex128)
        by_pulse = {}
        order = []
        for i, item in enumerate(H):
            if isinstance(item, list):
                l = next(ll for ll in range(L) if i in mapping[k][0][ll])
                if l not in by_pulse:
                    by_pulse[l] = np.zeros((N, N), dtype=np.complex128)
                    order.append(l)
                by_pulse[l] = by_pulse[l] + dense(item[0])
            else:
                drift = drift + dense(item)
        per_obj.append((drift, [(l, by_pulse[l]) for l in order]))
    M = 1 + max(len(terms) for _, terms in per_obj)
    cp.M = M
    ops = np.zeros((K, M, N, N), dtype=np.complex128)
    t2p = np.full((K, M), -2, dtype=np.int32)
    for k, (drift, terms) in enumerate(per_obj):
        ops[k, 0] = drift
        t2p[k, 0] = -1
        for m, (l, op) in enumerate(terms, start=1):
            ops[k, m] = op
            t2p[k, m] = l
    cp.term2pulse = t2p
    cp.op_norm = np.abs(ops).sum(axis=2).max(axis=2)  # 1-norm per term
    # most non-zero columns in a row of an objective's terms taken together
    # (diagonal included): sparse rows select the entries-in-registers sweep
    pattern = np.any(ops != 0.0, axis=1) | np.eye(N, dtype=bool)[None]
    cp.row_nnz = int(pattern.sum(axis=2).max())
    # real Hamiltonians: f*A is purely imaginary, kernels halve the multiplies
    cp.real_ops = bool(np.all(ops.imag == 0.0))
    ops_adj = np.conj(np.swapaxes(ops, 2, 3))

    # mu table
    mu_tab = np.zeros((K, max(L, 1), N, N), dtype=np.complex128)
    custom_mu = mu is not None and mu is not derivative_wrt_pulse
    for k, (drift, terms) in enumerate(per_obj):
        if custom_mu:
            for l in range(L):
                mu_tab[k, l] = _probe_mu(mu, objectives, k, pulses_for_mu,
                                         mapping, l, cp)
        else:
            for l, op in terms:
                mu_tab[k, l] = (1j * op) if is_super else op
    # (a custom mu with imaginary parts takes the problem out of the real family)
    cp.real_ops = cp.real_ops and bool(np.all(mu_tab.imag == 0.0))
    # sparse generators beyond the delta-polynomial family's reach (N > 16, e.g. the
    # 17-level transmon of notebook 05: tridiagonal drive, diagonal drift) also go
    # through the CSR kernels: a handful of non-zeros per row instead of N
    density = float(np.count_nonzero(ops)) / max(1, ops[..., 0, 0].size * N * N)
    if N > DENSE_NMAX or (N > SPARSE_NMIN and density <= 0.25):
        # large state vectors (Liouville space of notebook 06: N = 625): the
        # kernels take the matrices in CSR form (include/krotov_b200.h,
        # kq_sparse; numbering: terms | adjoint terms | mu)
        cp.sparse = _csr_bundle(
            [ops[k, m] for k in range(K) for m in range(M)]
            + [ops_adj[k, m] for k in range(K) for m in range(M)]
            + [mu_tab[k, l] for k in range(K) for l in range(max(L, 1))], N)
        # the most non-zeros one CTA works with: terms + mu / terms or adjoints
        per = np.diff(cp.sparse['mat_off'])
        Lm = max(L, 1)
        fw = per[:K * M].reshape(K, M).sum(axis=1)
        bw = per[K * M:2 * K * M].reshape(K, M).sum(axis=1)
        mu_n = per[2 * K * M:].reshape(K, Lm).sum(axis=1)
        cp.sparse['stage_nnz_update'] = int(np.max(fw + mu_n))
        cp.sparse['stage_nnz_prop'] = int(max(np.max(fw), np.max(bw)))
        cp.ops = cp.ops_adj = cp.mu = None
    else:
        # column-major storage: element (r,c) at c*N + r
        cp.ops = np.ascontiguousarray(np.swapaxes(ops, 2, 3))
        cp.ops_adj = np.ascontiguousarray(np.swapaxes(ops_adj, 2, 3))
        cp.mu = np.ascontiguousarray(np.swapaxes(mu_tab, 2, 3))

    cp.state_templates = [obj.initial_state for obj in objectives]
    cp.psi0 = np.array([cp.vec(obj.initial_state) for obj in objectives])
    targets = []
    for obj in objectives:
        t = obj.target
        try:
            ok = t is not None and not isinstance(t, str) and \
                dense(t).size == N
        except Exception:
            ok = False
        targets.append(cp.vec(t) if ok else None)
    cp.targets = (np.array(targets) if all(t is not None for t in targets)
                  else None)
    if any(hasattr(obj, 'weight') for obj in objectives):
        cp.weights = np.array(
            [float(getattr(obj, 'weight', 1.0)) for obj in objectives])
    return cp


DENSE_NMAX = 64   # largest state length the dense kernel families take
SPARSE_NMIN = 64  # above this, sparse generators (<= 25 % non-zeros) use the CSR family
# (measured: the 17-level transmon runs 3x slower through the CSR kernels than through the
# dense lane-per-row kernels -- every Horner step of the CSR kernel is a chain of dependent
# shared-memory lookups -- so the threshold equals DENSE_NMAX)


def _csr_bundle(mats, N):
    """CSR of a list of N x N matrices in the layout of ``kq_sparse``:
    row_ptr [n_mat, N+1] (relative to the matrix' offset), mat_off [n_mat+1],
    col, val; exact zeros are dropped."""
    import scipy.sparse
    row_ptr = np.zeros((len(mats), N + 1), dtype=np.int32)
    mat_off = np.zeros(len(mats) + 1, dtype=np.int64)
    cols, vals = [], []
    for i, a in enumerate(mats):
        c = scipy.sparse.csr_matrix(np.asarray(a, dtype=np.complex128))
        c.eliminate_zeros()
        c.sort_indices()
        row_ptr[i] = c.indptr
        mat_off[i + 1] = mat_off[i] + c.nnz
        cols.append(c.indices.astype(np.int32))
        vals.append(c.data.astype(np.complex128))
    col = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    val = np.concatenate(vals) if vals else np.zeros(0, np.complex128)
    if col.size == 0:      # keep the device arrays non-empty
        col, val = np.zeros(1, np.int32), np.zeros(1, np.complex128)
    out = dict(row_ptr=row_ptr, mat_off=mat_off, col=col, val=val,
               nnz=int(mat_off[-1]), col16=None, code16=None, dict=None)
    # dictionary coding: a Liouvillian repeats few distinct values (bit-exact
    # comparison), so a non-zero shrinks to a 16-bit column + a 16-bit code and
    # the matrices of one objective fit the shared memory of its CTA
    # Codes are numbered by first appearance in the order the kernel touches the
    # non-zeros -- matrix by matrix, the p-th entry of every row, rows (= lanes)
    # fastest -- so that neighbouring lanes mostly read neighbouring dictionary
    # entries (no shared-memory bank conflicts).
    nnz_tot = int(mat_off[-1])
    if nnz_tot > 0:
        mat_of = np.repeat(np.arange(len(mats)), np.diff(mat_off))
        row_of = np.concatenate([
            np.repeat(np.arange(N), np.diff(row_ptr[i])) for i in range(len(mats))])
        pos_of = np.arange(nnz_tot) - (mat_off[mat_of] + row_ptr[mat_of, row_of])
        order = np.lexsort((row_of, pos_of, mat_of))
        uniq, first, inv = np.unique(
            val[order].view(np.float64).reshape(-1, 2), axis=0,
            return_index=True, return_inverse=True)
        if len(uniq) <= 65536 and N <= 65536:
            rank = np.empty(len(uniq), dtype=np.int64)
            rank[np.argsort(first, kind='stable')] = np.arange(len(uniq))
            codes = np.empty(nnz_tot, dtype=np.int64)
            codes[order] = rank[inv.ravel()]
            table = np.empty((len(uniq), 2), dtype=np.float64)
            table[rank] = uniq
            out['dict'] = table.view(np.complex128).ravel()
            out['code16'] = codes.astype(np.uint16)
            out['col16'] = col.astype(np.uint16)
    return out


def _probe_mu(mu, objectives, k, pulses, mapping, l, cp):
    """Matrix of a user-supplied ``mu`` for (objective k, pulse l), obtained
    by evaluating it once (operator result) or by applying the returned
    callable to the basis states (linear map)."""
    N = cp.N

    def matrix_at(n):
        res = mu(objectives, k, pulses, mapping, l, n)
        try:
            a = dense(res)
            if a.shape == (N, N):
                return a
        except Exception:
            pass
        if not callable(res):
            raise NotImplementedError(
                "custom mu must return an operator or a callable")
        out = np.zeros((N, N), dtype=np.complex128)
        tmpl = cp.state_templates[k] if cp.state_templates else \
            objectives[k].initial_state
        for j in range(N):
            e = np.zeros(N, dtype=np.complex128)
            e[j] = 1
            img = res(cp.unvec(e, tmpl))
            out[:, j] = cp.vec(img)
        return out

    cp.state_templates = [o.initial_state for o in objectives]
    m0 = matrix_at(0)
    m1 = matrix_at(max(0, cp.NT - 1))
    if not np.allclose(m0, m1, rtol=1e-13, atol=1e-300):
        raise NotImplementedError(
            "time- or pulse-dependent mu cannot be lowered to the B200 "
            "sweep kernels")
    return m0
