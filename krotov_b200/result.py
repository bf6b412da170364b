"""Record of an optimisation (host-side bookkeeping, once per iteration).

Field names and semantics follow ``krotov.result.Result``
(/root/reference/src/krotov/result.py:18-278): ``optimized_controls`` hold
interval pulses while the optimisation runs and grid-point controls after
finalisation; ``all_pulses`` always hold interval pulses; ``iter_seconds``
are integer seconds (float seconds from CUDA events are kept in
``iter_seconds_device``).
"""
import logging
import pickle
import time
from datetime import datetime, timedelta

from .conversions import pulse_onto_tlist
from .objectives import Objective

__all__ = ['Result']


class Result:
    """Result of :func:`krotov_b200.optimize_pulses`."""

    time_fmt = "%Y-%m-%d %H:%M:%S"

    def __init__(self):
        self.objectives = []
        self.tlist = []
        self.iters = []
        self.iter_seconds = []
        self.iter_seconds_device = []
        self.info_vals = []
        self.tau_vals = []
        self.guess_controls = []
        self.optimized_controls = []
        self.controls_mapping = []
        self.all_pulses = []
        self.states = []
        self.start_local_time = None
        self.end_local_time = None
        self.message = ''

    def _fmt_time(self, t):
        return 'n/a' if t is None else time.strftime(self.time_fmt, t)

    @property
    def start_local_time_str(self):
        return self._fmt_time(self.start_local_time)

    @property
    def end_local_time_str(self):
        return self._fmt_time(self.end_local_time)

    def __str__(self):
        elapsed = ""
        try:
            t0 = datetime(*self.start_local_time[:6])
            t1 = datetime(*self.end_local_time[:6])
            elapsed = " (%s)" % timedelta(seconds=(t1 - t0).total_seconds())
        except Exception:
            pass
        lines = [
            "Krotov Optimization Result",
            "--------------------------",
            "- Started at %s" % self.start_local_time_str,
            "- Number of objectives: %d" % len(self.objectives),
            "- Number of iterations: %d" % (len(self.iters) - 1),
            "- Reason for termination: %s" % self.message,
            "- Ended at %s%s" % (self.end_local_time_str, elapsed),
        ]
        return "\n".join(lines)

    __repr__ = __str__

    def objectives_with_controls(self, controls):
        """Copy of :attr:`objectives` with `controls` (arrays on the grid
        points, or callables) substituted via :attr:`controls_mapping`.

        Raises:
            ValueError: wrong number of controls or wrong array length.
        """
        if len(controls) != len(self.guess_controls):
            raise ValueError("Expected %d controls, %d given"
                             % (len(self.guess_controls), len(controls)))
        for c in controls:
            try:
                if len(c) != len(self.tlist):
                    raise ValueError(
                        "controls are not defined on the points of the time "
                        "grid: control has %d values for %d time grid points"
                        % (len(c), len(self.tlist)))
            except TypeError:
                pass

        def plug(nested, mapping):
            out = [list(h) if isinstance(h, list) else h for h in nested]
            for c, positions in zip(controls, mapping):
                for i in positions:
                    out[i][1] = c
            return out

        out = []
        for k, obj in enumerate(self.objectives):
            m = self.controls_mapping[k]
            new = Objective(
                H=plug(obj.H, m[0]), initial_state=obj.initial_state,
                target=obj.target,
                c_ops=[plug(c, m[j + 1]) for j, c in enumerate(obj.c_ops)])
            out.append(new)
        return out

    @property
    def optimized_objectives(self):
        return self.objectives_with_controls(self.optimized_controls)

    def dump(self, filename):
        """Pickle the result; callable controls that cannot be pickled are
        dropped (``Objective.__getstate__``), as in the reference."""
        with open(filename, 'wb') as fh:
            pickle.dump(self, fh)

    @classmethod
    def load(cls, filename, objectives=None, finalize=False):
        """Load a :meth:`dump` file; pass `objectives` to restore callable
        controls, ``finalize=True`` to convert interval pulses of an
        unfinished optimisation to grid-point controls (result.py:190-245)."""
        log = logging.getLogger('krotov')
        with open(filename, 'rb') as fh:
            result = pickle.load(fh)
        if objectives is not None:
            result.objectives = objectives
        nt = len(result.tlist)
        for i, c in enumerate(result.optimized_controls):
            if len(c) == nt - 1:
                if finalize:
                    result.optimized_controls[i] = pulse_onto_tlist(c)
                else:
                    log.warning("Result.optimized_controls are not finalized."
                                " Consider loading with `finalize=True`.")
                    break
            elif len(c) != nt:
                log.error("Result.optimized_controls are incongruent with "
                          "Result.tlist")
                break
        return result
