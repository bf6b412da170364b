"""krotov_b200 -- B200-native engine for Krotov's method."""
from . import conversions, objectives, shapes, workloads  # noqa: F401
from .objectives import (Objective, ensemble_objectives,  # noqa: F401
                         gate_objectives, liouvillian)

__version__ = '0.1.0'
