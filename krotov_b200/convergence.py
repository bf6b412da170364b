"""Convergence checks: callables ``check(result) -> None | str`` evaluated
on the host once per iteration (reference: src/krotov/convergence.py).

The reference extracts the checked values with ``glom.glom(result, spec)``
(defaults ``('info_vals', T[-1])``, convergence.py:109, 167, 211-214).  glom is
an optional dependency here: if it is importable every `spec` is handed to it
unchanged; otherwise the subset the reference itself uses is interpreted by
:func:`extract` -- callables, attribute / key names (dotted strings), tuples
of such steps, and :data:`T` paths such as ``T[-1]`` or ``T.tau_vals[-1]``.
Anything else raises ``TypeError`` instead of being ignored.
"""
from operator import xor

try:   # pragma: no cover - not installed in the build image
    import glom as _glom
except ImportError:
    _glom = None

__all__ = ['Or', 'value_below', 'value_above', 'delta_below',
           'check_monotonic_error', 'check_monotonic_fidelity', 'dump_result',
           'T', 'extract']


class _Path:
    """Minimal stand-in for ``glom.T``: records item / attribute accesses and
    replays them on a target."""

    __slots__ = ('_ops',)

    def __init__(self, ops=()):
        object.__setattr__(self, '_ops', tuple(ops))

    def __getitem__(self, key):
        return _Path(self._ops + (('item', key),))

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Path(self._ops + (('attr', name),))

    def __repr__(self):
        out = 'T'
        for kind, key in self._ops:
            out += '[%r]' % (key,) if kind == 'item' else '.' + key
        return out

    def _apply(self, target):
        for kind, key in self._ops:
            target = target[key] if kind == 'item' else getattr(target, key)
        return target


T = _Path() if _glom is None else _glom.T


def _step(target, spec):
    if isinstance(spec, _Path):
        return spec._apply(target)
    if callable(spec):
        return spec(target)
    if isinstance(spec, str):
        for part in spec.split('.'):
            try:
                target = getattr(target, part)
            except AttributeError:
                try:
                    target = target[part]
                except (TypeError, KeyError, IndexError):
                    raise AttributeError(
                        "%r has no attribute or key %r"
                        % (type(target).__name__, part)) from None
        return target
    if isinstance(spec, int) and not isinstance(spec, bool):
        return target[spec]
    if isinstance(spec, tuple):
        for s in spec:
            target = _step(target, s)
        return target
    raise TypeError(
        "unsupported spec %r: without the glom package only callables, "
        "attribute names, integer indices, krotov_b200.convergence.T paths "
        "and tuples of those can be used" % (spec,))


def extract(result, spec, **kwargs):
    """``glom.glom(result, spec, **kwargs)``, or its built-in subset."""
    if _glom is not None:
        return _glom.glom(result, spec, **kwargs)
    if kwargs:
        raise TypeError("keyword arguments for glom need the glom package")
    return _step(result, spec)


_LOOKUP_ERRORS = (AttributeError, KeyError, IndexError) + (
    () if _glom is None else (_glom.GlomError,))


def Or(*funcs):
    """Logical Or of several checks: the result of the first one that
    evaluates to True (convergence.py:84-106)."""
    def check_convergence(result):
        for f in funcs:
            msg = f(result)
            if bool(msg) is True:
                return msg
        return None
    return check_convergence


def value_below(limit, spec=('info_vals', T[-1]), name=None, **kwargs):
    """Check ``value < limit`` for the value `spec` extracts from the Result
    (default: the last entry of ``info_vals``); `limit` may be a string such
    as '1e-3', which is then quoted verbatim in the message
    (convergence.py:109-164).  Lookup errors propagate, as in the reference
    (e.g. IndexError when no `info_hook` fills ``info_vals``)."""
    if name is None:
        name = str(spec)

    def check_convergence(result):
        v = extract(result, spec, **kwargs)
        if v < float(limit):
            return "%s < %s" % (name, limit)
        return None
    return check_convergence


def value_above(limit, spec=('info_vals', T[-1]), name=None, **kwargs):
    """Like :func:`value_below` for ``value > limit`` (convergence.py:167-208)."""
    if name is None:
        name = str(spec)

    def check_convergence(result):
        v = extract(result, spec, **kwargs)
        if v > float(limit):
            return "%s > %s" % (name, limit)
        return None
    return check_convergence


def delta_below(limit, spec1=('info_vals', T[-1]),
                spec0=('info_vals', T[-2]), absolute_value=True, name=None,
                **kwargs):
    """Check ``|v1 - v0| < limit`` (or ``v1 - v0 < limit``) for the values
    `spec1`, `spec0` extract (convergence.py:211-297).  A missing v0 XOR v1
    (first iteration) passes; if neither can be extracted the lookup error is
    re-raised."""
    if name is None:
        name = "Δ(%s,%s)" % (spec1, spec0)

    def check_convergence(result):
        pending = None
        vals = []
        for spec in (spec1, spec0):
            try:
                vals.append(extract(result, spec, **kwargs))
            except _LOOKUP_ERRORS as exc:
                vals.append(None)
                pending = exc
        v1, v0 = vals
        if xor(v1 is None, v0 is None):
            return None
        if pending is not None:
            raise pending
        delta = v1 - v0
        if absolute_value:
            delta = abs(delta)
        if delta < float(limit):
            return "%s < %s" % (name, limit)
        return None
    return check_convergence


_monotonic_error = delta_below(
    limit=0, spec1=('info_vals', T[-2]), spec0=('info_vals', T[-1]),
    absolute_value=False,
    name="Loss of monotonic convergence; error decrease")

_monotonic_fidelity = delta_below(
    limit=0, spec1=('info_vals', T[-1]), spec0=('info_vals', T[-2]),
    absolute_value=False,
    name="Loss of monotonic convergence; fidelity increase")


def check_monotonic_error(result):
    """Message if the last value of ``info_vals`` (an error) is larger than
    the one before (convergence.py:316-344)."""
    return _monotonic_error(result)


def check_monotonic_fidelity(result):
    """Message if the last value of ``info_vals`` (a fidelity) is smaller than
    the one before (convergence.py:347-367)."""
    return _monotonic_fidelity(result)


def dump_result(filename, every=10):
    """Check that never converges but dumps the result every `every`
    iterations to ``filename.format(iter=...)`` (convergence.py:370-419)."""
    every = int(every)
    if every <= 0:
        raise ValueError("every must be > 0")

    def check_convergence(result):
        iteration = result.iters[-1]
        if iteration % every == 0:
            result.dump(filename.format(iter=iteration))
        return None
    return check_convergence
