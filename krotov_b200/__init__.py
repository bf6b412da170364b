"""krotov_b200 -- B200-native engine for Krotov's method of optimal control.

Drop-in for the hot path of ``qucontrol/krotov``: the same
``optimize_pulses`` / ``Objective`` / ``propagator`` / ``chi_constructor`` /
``mu`` plugin surface, with the backward-propagate / pulse-update /
forward-propagate sweeps executed by hand-written sm_100a CUDA kernels
(``csrc/``) behind a C ABI (``include/krotov_b200.h``).  Use as::

    import krotov_b200 as krotov

There is no CPU fallback: without the built library and a CUDA device
``optimize_pulses`` raises :class:`EngineUnavailable`.  See DESIGN.md.
"""
from . import (conversions, convergence, functionals, info_hooks, mu,  # noqa
               objectives, perfect_entanglers, propagators, result,
               second_order, shapes, workloads)
from ._lib import EngineUnavailable, KqError  # noqa: F401
from .objectives import (Objective, ensemble_objectives,  # noqa: F401
                         gate_objectives, liouvillian)
from .optimize import optimize_pulses  # noqa: F401
from .result import Result  # noqa: F401

__version__ = '0.1.0'

__all__ = [
    'Objective', 'Result', 'conversions', 'convergence', 'ensemble_objectives',
    'functionals', 'gate_objectives', 'info_hooks', 'liouvillian', 'mu',
    'objectives', 'optimize_pulses', 'perfect_entanglers', 'propagators',
    'result', 'second_order',
    'shapes', 'workloads', 'EngineUnavailable', 'KqError',
]
