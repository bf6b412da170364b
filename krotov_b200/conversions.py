"""Structural conversions between the QuTiP nested-list problem description
and the dense pulse layout the sm_100a sweep kernels consume.

Public names and results follow the reference's ``krotov.conversions``
(/root/reference/src/krotov/conversions.py).  Two things defined here are
parity-critical for the hot path:

* controls live on the ``nt`` grid *points*, pulses on the ``nt-1``
  *intervals*; the two are related by the averaging recurrences of
  :func:`control_onto_interval` / :func:`pulse_onto_tlist`, which are
  evaluated sequentially in float64 exactly as the reference does
  (conversions.py:333-390), because the guess pulse is an input to every
  kernel;
* controls are identified by *object identity* (conversions.py:43-58,174),
  which is what couples an ensemble of objectives through one shared pulse.
"""
import logging
import warnings

import numpy as np

__all__ = [
    'control_onto_interval',
    'discretize',
    'extract_controls',
    'extract_controls_mapping',
    'plug_in_pulse_values',
    'pulse_onto_tlist',
    'pulse_options_dict_to_list',
]


def _is_term(item):
    return isinstance(item, list)


def _index_by_identity(control, controls):
    """Index of `control` in `controls`; arrays compare by identity, other
    objects by ``==`` (conversions.py:43-58)."""
    if isinstance(control, np.ndarray):
        for i, c in enumerate(controls):
            if c is control:
                return i
        return -1
    try:
        return controls.index(control)
    except ValueError:
        return -1


def _sample(control, points, args, kwargs):
    # float() of a complex value raises TypeError; numpy complex scalars emit
    # ComplexWarning, which we escalate, so complex controls are rejected
    # (conversions.py:103,122-126).
    with warnings.catch_warnings():
        warnings.simplefilter("error", category=np.exceptions.ComplexWarning)
        return np.array(
            [float(control(t, *args, **kwargs)) for t in points],
            dtype=np.float64,
        )


def discretize(control, tlist, args=(None,), kwargs=None, via_midpoints=False):
    """Sample `control` on `tlist` (array of ``nt`` float64 values).

    With ``via_midpoints=True`` a callable is sampled at the interval
    midpoints -- first and last sample pinned to ``tlist[0]`` / ``tlist[-1]``,
    midpoint offset taken from the *first* interval as in the reference
    (conversions.py:108-119) -- and mapped back to the grid points with
    :func:`pulse_onto_tlist`, so that converting back to intervals is exact.

    Raises:
        TypeError: control is neither callable nor array-like (or is complex).
        ValueError: array control whose length differs from ``len(tlist)``.
    """
    if callable(control):
        if kwargs is None:
            kwargs = {}
        if via_midpoints:
            mid = (tlist + 0.5 * (tlist[1] - tlist[0]))[:-1]
            mid[0] = tlist[0]
            mid[-1] = tlist[-1]
            return pulse_onto_tlist(_sample(control, mid, args, kwargs))
        return _sample(control, tlist, args, kwargs)
    if isinstance(control, (np.ndarray, list)):
        with warnings.catch_warnings():
            warnings.simplefilter(
                "error", category=np.exceptions.ComplexWarning
            )
            values = np.array([float(v) for v in control], dtype=np.float64)
        if len(values) != len(tlist):
            raise ValueError(
                "If control is an array, it must of the same length as tlist"
            )
        return values
    raise TypeError(
        "control must be either a callable func(t, args) or a numpy array"
    )


def extract_controls(objectives):
    """Unique controls (by identity) in the ``H`` of all `objectives`, in
    order of first appearance (conversions.py:140-166)."""
    controls = []
    for objective in objectives:
        for item in objective.H:
            if _is_term(item):
                assert len(item) == 2
                if _index_by_identity(item[1], controls) < 0:
                    controls.append(item[1])
    return controls


def _positions(nested_list, control):
    return [
        i
        for i, item in enumerate(nested_list)
        if _is_term(item) and len(item) == 2 and item[1] is control
    ]


def extract_controls_mapping(objectives, controls):
    """``mapping[k][0][l]`` = indices of ``objectives[k].H`` driven by control
    ``l``; ``mapping[k][1+ic][l]`` the same for collapse operator ``ic``
    (conversions.py:179-254)."""
    mapping = []
    for objective in objectives:
        per_objective = [[_positions(objective.H, c) for c in controls]]
        for c_op in objective.c_ops:
            per_objective.append([_positions(c_op, c) for c in controls])
        mapping.append(per_objective)
    return mapping


def pulse_options_dict_to_list(pulse_options, controls):
    """Options dicts in the order of `controls`; array controls are looked up
    by ``id(control)`` (conversions.py:257-285).

    Raises:
        ValueError: a control has no entry in `pulse_options`.
    """
    if len(pulse_options) > len(controls):
        logging.getLogger('krotov').warning(
            "pulse_options contains extra elements that are not in `controls`"
        )
    out = []
    for control in controls:
        try:
            try:
                out.append(pulse_options[control])
            except TypeError:  # unhashable (numpy array)
                out.append(pulse_options[id(control)])
        except KeyError:
            raise ValueError(
                "The control %s does not have any associated pulse options"
                % str(control)
            )
    return out


def plug_in_pulse_values(H, pulses, mapping, time_index, conjugate=False):
    """Copy of nested list `H` with each control replaced by the scalar
    ``pulses[l][time_index]`` (conversions.py:288-330).

    The CUDA sweeps never call this: the same substitution is done in-kernel
    from the ``term2pulse`` table.  It is kept for step-wise host use
    (e.g. :meth:`Objective.propagate`-style loops and tests).
    """
    if isinstance(H, list):
        H = [list(item) if _is_term(item) else item for item in H]
    for pulse, positions in zip(pulses, mapping):
        for i in positions:
            value = pulse[time_index]
            H[i][1] = np.conjugate(value) if conjugate else value
    return H


def control_onto_interval(control):
    """Grid-point values (``nt``) → interval values (``nt-1``).

    First/last interval take the first/last grid value; in between
    ``pulse[i] = 2*control[i] - pulse[i-1]`` (sequential float64 recurrence,
    conversions.py:333-364).
    """
    if not isinstance(control, np.ndarray):
        raise ValueError(
            "Not implemented: control type %s" % control.__class__.__name__
        )
    assert control.ndim == 1
    n = len(control) - 1
    pulse = np.zeros(n, dtype=control.dtype.type)
    pulse[0] = control[0]
    for i in range(1, n):
        pulse[i] = 2.0 * control[i] - pulse[i - 1]
    pulse[-1] = control[-1]
    return pulse


def pulse_onto_tlist(pulse):
    """Interval values (``nt-1``) → grid-point values (``nt``): end points
    copied, interior points the mean of the two adjacent intervals
    (conversions.py:368-390)."""
    n = len(pulse) + 1
    control = np.zeros(n, dtype=pulse.dtype.type)
    control[0] = pulse[0]
    if n > 2:
        control[1:-1] = 0.5 * (pulse[:-1] + pulse[1:])
    control[-1] = pulse[-1]
    return control
