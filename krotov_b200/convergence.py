"""Convergence checks: callables ``check(result) -> None | str`` evaluated
on the host once per iteration (reference: src/krotov/convergence.py)."""
import numpy as np

__all__ = ['Or', 'value_below', 'value_above', 'delta_below',
           'check_monotonic_error', 'check_monotonic_fidelity', 'dump_result']


def _last(result, index=-1, attr='info_vals'):
    vals = getattr(result, attr)
    v = vals[index]
    return v[0] if isinstance(v, (tuple, list)) else v


def Or(*funcs):
    """First non-None message of the given checks (convergence.py:84-106)."""
    def check(result):
        for f in funcs:
            msg = f(result)
            if msg is not None:
                return msg
        return None
    return check


def value_below(limit, spec=None, name=None, **kwargs):
    """Converged when the last info value (or ``spec(result)``) is below
    `limit` (float or string like '1e-3')."""
    lim = float(limit)
    label = name or 'value'

    def check(result):
        try:
            v = spec(result) if callable(spec) else _last(result)
        except (IndexError, TypeError, AttributeError):
            return None
        if v is not None and v < lim:
            return "%s < %s" % (label, limit)
        return None
    return check


def value_above(limit, spec=None, name=None, **kwargs):
    lim = float(limit)
    label = name or 'value'

    def check(result):
        try:
            v = spec(result) if callable(spec) else _last(result)
        except (IndexError, TypeError, AttributeError):
            return None
        if v is not None and v > lim:
            return "%s > %s" % (label, limit)
        return None
    return check


def delta_below(limit, spec1=None, spec0=None, absolute_value=True,
                name=None, **kwargs):
    """Converged when the change between the last two info values is below
    `limit` (convergence.py:211-297)."""
    lim = float(limit)
    label = name or ('Δvalue' if True else '')

    def check(result):
        try:
            v1 = spec1(result) if callable(spec1) else _last(result, -1)
            v0 = spec0(result) if callable(spec0) else _last(result, -2)
        except (IndexError, TypeError, AttributeError):
            return None
        delta = v1 - v0
        if absolute_value:
            delta = abs(delta)
        if delta < lim:
            return "%s < %s" % (label, limit)
        return None
    return check


def check_monotonic_error(result):
    """Message if the error (last info value) increased
    (convergence.py:316-341)."""
    try:
        if _last(result, -1) - _last(result, -2) > 0:
            return "Loss of monotonic convergence; error decrease < 0"
    except (IndexError, TypeError):
        pass
    return None


def check_monotonic_fidelity(result):
    try:
        if _last(result, -2) - _last(result, -1) > 0:
            return "Loss of monotonic convergence; fidelity increase < 0"
    except (IndexError, TypeError):
        pass
    return None


def dump_result(filename, every=10):
    """Check that never converges but dumps the result every `every`
    iterations to ``filename.format(iter=...)`` (convergence.py:370-419)."""
    def check(result):
        it = result.iters[-1]
        if it % every == 0:
            result.dump(filename.format(iter=it))
        return None
    return check
