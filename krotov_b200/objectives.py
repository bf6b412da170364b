"""Problem description: :class:`Objective` and the constructors for gate and
ensemble optimisations.

These are the *inputs* of the hot path.  Names, argument
meaning and error behaviour follow the reference's ``krotov.objectives``
(/root/reference/src/krotov/objectives.py:96-258, 704-1121); the arithmetic
is done on dense numpy data (or on the user's own Qobj-like objects through
their operators), never inside a CUDA kernel -- the problem compiler
(:mod:`krotov_b200.compiler`) lowers the result to device tensors once per
``optimize_pulses`` call.
"""
import copy
import itertools

import numpy as np

from ._dense import adjoint_of, dense, is_quantum_object, kind_of

__all__ = [
    'Objective',
    'gate_objectives',
    'ensemble_objectives',
    'liouvillian',
]


def _copy_nested(item):
    """Copy the list structure of a nested-list operator, sharing leaves."""
    if isinstance(item, list):
        return [list(h) if isinstance(h, list) else h for h in item]
    return item


def _adjoint(op, ignore_errors=False):
    """Adjoint of an operator/state or of every operator in a nested list;
    controls are left untouched (objectives.py:51-93 in the reference)."""
    if isinstance(op, list):
        out = []
        for item in op:
            if isinstance(item, list):
                if len(item) != 2:
                    if ignore_errors:
                        return op
                    raise ValueError(
                        "%s is not the in the expected format of the "
                        "two-element list '[operator, control]'" % item
                    )
                out.append([_adjoint(item[0]), item[1]])
            else:
                out.append(_adjoint(item))
        return out
    if op is None or isinstance(op, str):
        return op
    try:
        return adjoint_of(op)
    except AttributeError:
        if ignore_errors:
            return op
        raise ValueError("Cannot calculate adjoint of %s" % op)


def _same(a, b):
    """Structural equality that understands nested lists and arrays."""
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        try:
            return bool(np.array_equal(a, b))
        except Exception:
            return False
    try:
        return bool(a == b)
    except Exception:
        return a is b


class Objective:
    """One control objective: steer `initial_state` towards `target` under
    the (time-dependent) generator `H`.

    Args:
        initial_state: ket or density matrix (Qobj-like or ndarray).
        H: operator, or nested list ``[H0, [H1, control], ...]`` in QuTiP's
            time-dependent format; a control is a callable ``f(t, args)`` or
            an array on the time grid.  May be a Liouvillian.
        target: target state, or any object a custom `chi_constructor`
            understands (e.g. the string ``'PE'``).
        c_ops: list of collapse operators (only meaningful for propagators
            that use them; prefer a Liouvillian `H`, see :func:`liouvillian`).

    Raises:
        ValueError: invalid argument types, unless the class attribute
            :attr:`type_checking` is False.
    """

    _default_attribs = ['initial_state', 'H', 'target', 'c_ops']

    str_use_unicode = True
    type_checking = True
    """If False, skip the argument checks in the constructor (same switch as
    in the reference, objectives.py:154-158)."""

    def __init__(self, *, initial_state, H, target, c_ops=None):
        if c_ops is None:
            c_ops = []
        if self.type_checking:
            if not (isinstance(H, list) or is_quantum_object(H)):
                raise ValueError(
                    "Invalid H, must be a Qobj, or a nested list, not %s"
                    % H.__class__.__name__
                )
            if not is_quantum_object(initial_state):
                raise ValueError(
                    "Invalid initial_state: must be Qobj, not %s"
                    % initial_state.__class__.__name__
                )
            if not isinstance(c_ops, list):
                raise ValueError(
                    "Invalid c_ops: must be a list, not %s"
                    % c_ops.__class__.__name__
                )
        self.H = H
        self.initial_state = initial_state
        self.target = target
        self.c_ops = c_ops

    def _extra_attribs(self):
        return [a for a in self.__dict__ if a not in self._default_attribs]

    def __copy__(self):
        # list structure by value, operators/controls by reference
        new = Objective(
            H=_copy_nested(self.H),
            initial_state=self.initial_state,
            target=self.target,
            c_ops=[_copy_nested(c) for c in self.c_ops],
        )
        for attr in self._extra_attribs():
            setattr(new, attr, getattr(self, attr))
        return new

    def __deepcopy__(self, memo):
        new = Objective(
            H=copy.deepcopy(self.H, memo),
            initial_state=copy.deepcopy(self.initial_state, memo),
            target=copy.deepcopy(self.target, memo),
            c_ops=[copy.deepcopy(c, memo) for c in self.c_ops],
        )
        for attr in self._extra_attribs():
            setattr(new, attr, copy.deepcopy(getattr(self, attr), memo))
        return new

    def __eq__(self, other):
        if other.__class__ is not self.__class__:
            return NotImplemented
        if self.__dict__.keys() != other.__dict__.keys():
            return False
        return all(
            _same(getattr(self, a), getattr(other, a)) for a in self.__dict__
        )

    def __ne__(self, other):
        eq = self.__eq__(other)
        return eq if eq is NotImplemented else not eq

    __hash__ = None

    def adjoint(self):
        """Objective with every operator and state replaced by its adjoint;
        controls (assumed real) and non-state targets unchanged.  The
        backward sweep runs under ``adjoint().H`` (optimize.py:263,863)."""
        adj = Objective(
            H=_adjoint(self.H),
            initial_state=_adjoint(self.initial_state),
            target=_adjoint(self.target, ignore_errors=True),
            c_ops=[_adjoint(op) for op in self.c_ops],
        )
        for attr in self._extra_attribs():
            setattr(adj, attr, getattr(self, attr))
        return adj

    def summarize(self, use_unicode=True, reset_symbol_counters=False):
        """One-line description (dimensions and number of control terms)."""
        def _dim(x):
            try:
                return "x".join(str(s) for s in dense(x).shape)
            except Exception:
                return str(x)
        H = self.H if isinstance(self.H, list) else [self.H]
        n_ctrl = sum(1 for h in H if isinstance(h, list))
        arrow = '→' if (use_unicode and self.str_use_unicode) else 'to'
        return "state[%s] %s target[%s] via %d drift + %d control term(s)" % (
            _dim(self.initial_state), arrow, _dim(self.target),
            len(H) - n_ctrl, n_ctrl,
        )

    def __str__(self):
        return self.summarize()

    def __repr__(self):
        return "%s[%s]" % (self.__class__.__name__, self.summarize())

    def __getstate__(self):
        # callables that cannot be pickled (lambdas) are replaced by None,
        # like the reference's control placeholders (objectives.py:581-636)
        import pickle

        def _strip(x):
            if isinstance(x, list):
                return [_strip(v) for v in x]
            if callable(x) and not is_quantum_object(x):
                try:
                    pickle.dumps(x)
                    return x
                except Exception:
                    return None
            return x

        state = dict(self.__dict__)
        state['H'] = _strip(state['H'])
        state['c_ops'] = _strip(state['c_ops'])
        return state

    def propagate(self, tlist, *, propagator, rho0=None, H=None, c_ops=None,
                  e_ops=None, args=None, expect=None):
        """Propagate step by step over `tlist` with a host `propagator`
        (piecewise-constant controls on the *intervals*; objectives.py:338-426
        in the reference).  Returns an object with ``times``, ``states`` and
        ``expect`` like ``qutip.solver.Result``.

        This is an analysis helper outside the accelerated path: it calls
        `propagator` once per interval on the host."""
        from .conversions import (control_onto_interval, discretize,
                                  extract_controls, extract_controls_mapping,
                                  plug_in_pulse_values)
        H = self.H if H is None else H
        c_ops = self.c_ops if c_ops is None else c_ops
        e_ops = [] if e_ops is None else e_ops
        args = {} if args is None else args
        if expect is None:
            def expect(op, state):
                s, o = dense(state), dense(op)
                if s.shape[1] == 1:
                    return complex((s.conj().T @ o @ s)[0, 0])
                return complex(np.trace(o @ s))

        class _Result:
            pass

        result = _Result()
        result.solver = getattr(propagator, '__name__',
                                propagator.__class__.__name__)
        result.times = np.array(tlist)
        result.states = []
        result.expect = [[] for _ in e_ops]
        result.num_expect = len(e_ops)
        result.num_collapse = len(c_ops)
        state = self.initial_state if rho0 is None else rho0

        def _record(state):
            if e_ops:
                for i, op in enumerate(e_ops):
                    result.expect[i].append(expect(op, state))
            else:
                result.states.append(state)

        _record(state)
        controls = extract_controls([self])
        mapping = extract_controls_mapping([self], controls)[0]
        pulses = [
            control_onto_interval(discretize(c, tlist, args=(args,)))
            for c in controls
        ]
        for n in range(len(tlist) - 1):
            H_n = plug_in_pulse_values(H, pulses, mapping[0], n)
            c_n = [
                plug_in_pulse_values(c, pulses, mapping[ic + 1], n)
                for ic, c in enumerate(c_ops)
            ]
            state = propagator(H_n, state, tlist[n + 1] - tlist[n], c_n,
                               initialize=True)
            _record(state)
        result.expect = [np.array(a) for a in result.expect]
        return result


# ---------------------------------------------------------------------------
# constructors


def _outer(a, b):
    """|a><b| for Qobj-likes or column-vector arrays."""
    if hasattr(a, 'dag') and not isinstance(a, np.ndarray):
        return a * b.dag()
    return dense(a) @ dense(b).conj().T


def _rho1(basis):
    d = len(basis)
    return sum(
        (2 * (d - i) / (d * (d + 1))) * _outer(psi, psi)
        for i, psi in enumerate(basis)
    )


def _rho2(basis):
    d = len(basis)
    return (1.0 / d) * sum(
        _outer(a, b) for a, b in itertools.product(basis, repeat=2)
    )


def _rho3(basis):
    d = len(basis)
    return (1.0 / d) * sum(_outer(psi, psi) for psi in basis)


def _bell_objectives(basis_states, target, H, c_ops):
    """Bell-basis objectives for perfect-entangler / local-invariant
    optimisation (Makhlin's "Theorem 1" basis; objectives.py:1035-1051)."""
    if len(basis_states) != 4:
        raise ValueError(
            "Optimization towards a two-qubit gate requires 4 basis_states"
        )
    b = basis_states
    r = np.sqrt(2)
    bell = [
        (b[0] + b[3]) / r,
        (1j * b[1] + 1j * b[2]) / r,
        (b[1] - b[2]) / r,
        (1j * b[0] - 1j * b[3]) / r,
    ]
    return [
        Objective(initial_state=psi, target=target, H=H, c_ops=c_ops)
        for psi in bell
    ]


def gate_objectives(basis_states, gate, H, *, c_ops=None,
                    local_invariants=False, liouville_states_set=None,
                    weights=None, normalize_weights=True):
    """Objectives for optimising towards the quantum gate `gate`.

    For an ``n x n`` `gate` and ``n`` `basis_states`, objective ``j`` maps
    ``basis_states[j]`` to ``sum_i gate[i, j] basis_states[i]``.  With
    `liouville_states_set` in ``'full' | '3states' | 'd+1'`` the objectives
    are the density-matrix sets of Goerz et al., NJP 16, 055012 (2014).
    ``gate='PE'`` (or ``local_invariants=True`` with a 4x4 gate) yields the
    four Bell-basis objectives whose target is the string ``'PE'`` (resp. the
    gate).  `weights` are attached as ``objective.weight`` (normalised to
    sum to the number of objectives unless `normalize_weights` is False);
    zero-weight objectives are dropped.  Semantics and errors as
    objectives.py:704-1032 of the reference.
    """
    if isinstance(gate, str):
        if gate.lower().replace(' ', '_') in ('pe', 'perfect_entangler'):
            return _bell_objectives(basis_states, 'PE', H, c_ops)
        raise ValueError(
            "gate must be either a square matrix, or one of the strings "
            "'PE' or 'perfect_entangler', not '" + gate + "'"
        )
    if local_invariants:
        if not gate.shape == (4, 4):
            raise ValueError(
                "If local_invariants is True, gate must be a 4 × 4 matrix, "
                "not " + str(gate.shape)
            )
        return _bell_objectives(basis_states, gate, H, c_ops)
    n = len(basis_states)
    if not gate.shape[0] == gate.shape[1] == n:
        raise ValueError(
            "gate must be a matrix of the same dimension as the number of "
            "basis states"
        )
    mapped = [
        sum(complex(gate[i, j]) * basis_states[i] for i in range(n))
        for j in range(n)
    ]
    # permutation-like gates: reuse the identical basis-state objects
    for j, state in enumerate(mapped):
        for basis_state in basis_states:
            if _same(state, basis_state):
                mapped[j] = basis_state
    if liouville_states_set is None:
        initial, targets = list(basis_states), mapped
    else:
        key = liouville_states_set.replace(" ", "").lower()
        if key == 'full':
            initial = [_outer(a, b) for a, b
                       in itertools.product(basis_states, repeat=2)]
            targets = [_outer(a, b) for a, b
                       in itertools.product(mapped, repeat=2)]
        elif key == '3states':
            initial = [f(basis_states) for f in (_rho1, _rho2, _rho3)]
            targets = [f(mapped) for f in (_rho1, _rho2, _rho3)]
        elif key == 'd+1':
            initial = [_outer(p, p) for p in basis_states]
            initial.append(_rho2(basis_states))
            targets = [_outer(p, p) for p in mapped]
            targets.append(_rho2(mapped))
        else:
            raise ValueError(
                "Invalid `liouville_states_set`: %s" % liouville_states_set
            )
    objectives = [
        Objective(initial_state=s, target=t, H=H, c_ops=c_ops)
        for s, t in zip(initial, targets)
    ]
    if weights is not None:
        if len(weights) != len(objectives):
            raise ValueError(
                "If weight are given, there must be a weight for each "
                "objective"
            )
        if normalize_weights:
            weights = len(objectives) * np.array(weights) / np.sum(weights)
        for i in reversed(range(len(objectives))):
            w = float(weights[i])
            if w < 0:
                raise ValueError("weights must be greater than zero")
            objectives[i].weight = w
            if w == 0:
                del objectives[i]
    return objectives


def ensemble_objectives(objectives, Hs, *, keep_original_objectives=True):
    """Objectives for an ensemble (robustness) optimisation: one copy of every
    objective per generator in `Hs`, sharing the original control objects so
    that all copies are driven by the same pulses (objectives.py:1054-1094)."""
    out = list(objectives) if keep_original_objectives else []
    for H in Hs:
        for obj in objectives:
            out.append(
                Objective(H=H, initial_state=obj.initial_state,
                          target=obj.target, c_ops=obj.c_ops)
            )
    return out


def _dense_liouvillian(H, c_ops):
    """Column-stacking Lindblad superoperator of dense `H`, `c_ops`:
    ``-i(I⊗H − Hᵀ⊗I) + Σ_C [C*⊗C − ½ I⊗C†C − ½ (C†C)ᵀ⊗I]``."""
    first = H if H is not None else c_ops[0]
    d = dense(first).shape[0]
    eye = np.eye(d)
    L = np.zeros((d * d, d * d), dtype=np.complex128)
    if H is not None:
        h = dense(H)
        L += -1j * (np.kron(eye, h) - np.kron(h.T, eye))
    for c in c_ops:
        C = dense(c)
        cdc = C.conj().T @ C
        L += np.kron(C.conj(), C)
        L += -0.5 * np.kron(eye, cdc)
        L += -0.5 * np.kron(cdc.T, eye)
    return L


def _liouvillian_one(H, c_ops):
    if not isinstance(H, np.ndarray) and hasattr(H, 'full'):
        try:  # real QuTiP available: keep Qobj dims/superrep
            import qutip
            return qutip.liouvillian(H, c_ops)
        except ImportError:
            pass
    return _dense_liouvillian(H, list(c_ops))


def liouvillian(H, c_ops):
    """Liouvillian (super-operator, column-stacking ``vec``) of Hamiltonian
    `H` -- an operator or a nested list with a drift term -- and constant
    Lindblad operators `c_ops`.  The dissipators are attached to the drift
    term; control terms become ``[-i[H_l, .], control]``
    (objectives.py:1097-1121).  Array inputs give array outputs."""
    if isinstance(H, list):
        out = []
        pending = list(c_ops)
        for spec in H:
            if isinstance(spec, list):
                out.append([_liouvillian_one(spec[0], []), spec[1]])
            else:
                out.append(_liouvillian_one(spec, pending))
                pending = []
        assert len(pending) == 0, "No drift Hamiltonian"
        return out
    if is_quantum_object(H):
        return _liouvillian_one(H, c_ops)
    raise ValueError(
        "H must either be a Qobj, or a time-dependent Hamiltonian in "
        "nested-list format"
    )
