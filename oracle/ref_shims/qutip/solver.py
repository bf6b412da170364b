"""Shim of ``qutip.solver`` names imported at
/root/reference/src/krotov/objectives.py:18-19 (test infrastructure only)."""


class Options:
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class Result:
    def __init__(self):
        self.solver = None
        self.times = None
        self.states = []
        self.expect = []
        self.num_expect = 0
        self.num_collapse = 0
        self.ntraj = None
        self.seeds = None
        self.col_times = None
        self.col_which = None
