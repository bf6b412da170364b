"""``import krotov`` -> the B200 engine.

Put this directory in front of the reference on ``sys.path`` (for example
``PYTHONPATH=/path/to/repo/compat:/path/to/repo``) and the reference's scripts and
notebooks run against ``krotov_b200`` unchanged: the package name, the submodule
names and the names they export are the reference's
(/root/reference/src/krotov/__init__.py:40-65).  It lives outside the repository
root on purpose -- the test infrastructure imports the real reference under the
same name (oracle/make_golden.py, oracle/run_reference.py) and must not find
this alias first."""
import importlib
import sys

import krotov_b200 as _engine
from krotov_b200 import *  # noqa: F401,F403
from krotov_b200 import __all__ as _all

__version__ = _engine.__version__
__all__ = list(_all)

_SUBMODULES = ('convergence', 'conversions', 'functionals', 'info_hooks', 'mu',
               'objectives', 'optimize', 'parallelization', 'propagators',
               'result', 'second_order', 'shapes')
for _name in _SUBMODULES:
    _mod = importlib.import_module('krotov_b200.' + _name)
    sys.modules[__name__ + '.' + _name] = _mod
    if _name != 'optimize':          # `krotov.optimize_pulses` is the function
        globals()[_name] = _mod
optimize_pulses = _engine.optimize_pulses
