"""Update-shape and guess-pulse envelope functions.

Same public names, argument meaning and values as the reference's
``krotov.shapes`` (/root/reference/src/krotov/shapes.py:20-174).  They are
evaluated on the host once per :func:`krotov_b200.optimize_pulses` call by
:func:`krotov_b200.conversions.discretize`; the sampled arrays are what the
CUDA sweep kernels consume.
"""
import functools
import math

import numpy as np

__all__ = [
    'qutip_callback',
    'zero_shape',
    'one_shape',
    'flattop',
    'box',
    'blackman',
]

_TWO_PI = 2.0 * np.pi
_FOUR_PI = 4.0 * np.pi


def qutip_callback(func, **kwargs):
    """Turn ``func(t, **params)`` into a QuTiP-style control ``f(t, args)``.

    Parameters fixed in `kwargs` are bound now; the remaining ones are taken
    from the `args` dict at call time (shapes.py:20-38 in the reference).
    """
    bound = functools.partial(func, **kwargs)

    def callback(t, args):
        return bound(t, **({} if args is None else args))

    return callback


def zero_shape(t):
    """S(t) = 0 (disables the update of a control)."""
    return 0


def one_shape(t):
    """S(t) = 1."""
    return 1


def box(t, t_start, t_stop):
    """1 on the closed interval [t_start, t_stop], 0 outside."""
    return 1.0 if t_start <= t <= t_stop else 0.0


_box_vec = np.vectorize(box)


def blackman(t, t_start, t_stop, a=0.16):
    r"""Blackman window :math:`\frac12(1-a-\cos(2\pi x)+a\cos(4\pi x))`,
    :math:`x=(t-t_0)/(t_1-t_0)`, zero outside ``[t_start, t_stop]``.

    Accepts a scalar or an array `t` (shapes.py:131-174 in the reference; the
    expression is evaluated in the same order so sampled guess pulses agree
    bit for bit).
    """
    T = t_stop - t_start
    window = (
        1.0
        - a
        - np.cos(_TWO_PI * (t - t_start) / T)
        + a * np.cos(_FOUR_PI * (t - t_start) / T)
    )
    # (scalar fast path: np.vectorize costs ~25 us per call, and guess pulses are sampled
    # point by point -- the product is the same IEEE operation either way)
    b = box(t, t_start, t_stop) if np.ndim(t) == 0 else _box_vec(t, t_start, t_stop)
    return 0.5 * b * window


def flattop(t, t_start, t_stop, t_rise, t_fall=None, func='blackman'):
    """Flat-top envelope: 0 → 1 over `t_rise`, 1 → 0 over `t_fall`.

    ``func`` selects the ramp: half a Blackman window or a sine-squared edge
    (shapes.py:51-107 in the reference).
    """
    if t_fall is None:
        t_fall = t_rise
    # (branches written out: guess pulses and update shapes are sampled point by
    # point, a pair of closures per call showed up in the set-up time)
    if func == 'blackman':
        if not (t_start <= t <= t_stop):
            return 0.0
        if t <= t_start + t_rise:
            return blackman(t, t_start, t_start + 2 * t_rise)
        if t >= t_stop - t_fall:
            return blackman(t, t_stop - 2 * t_fall, t_stop)
        return 1.0
    if func == 'sinsq':
        if not (t_start <= t <= t_stop):
            return 0.0
        if t <= t_start + t_rise:
            return np.sin(np.pi * (t - t_start) / (2.0 * t_rise)) ** 2
        if t >= t_stop - t_fall:
            return np.sin(np.pi * (t - t_stop) / (2.0 * t_fall)) ** 2
        return 1.0
    raise ValueError("Invalid func: %s" % func)
