"""Derivative of the generator with respect to a pulse, dH/d eps_l.

Same signature and result as ``krotov.mu.derivative_wrt_pulse``
(/root/reference/src/krotov/mu.py:74-140).  Inside
:func:`krotov_b200.optimize_pulses` the default ``mu`` is *lowered*: it is
constant for controls that enter linearly, so the problem compiler evaluates
it once into the ``mu[k][l]`` table the fused sweep kernel reads, instead of
rebuilding it K*L*(nt-1) times per iteration as the reference loop does
(optimize.py:457-464).  This host version serves user code and custom hooks.
"""
from ._dense import kind_of

__all__ = ['derivative_wrt_pulse']


def derivative_wrt_pulse(objectives, i_objective, pulses, pulses_mapping,
                         i_pulse, time_index):
    """Operator (or zero map) representing dH/d eps for objective
    `i_objective` and pulse `i_pulse`; `time_index` is unused because the
    standard equations of motion are linear in the controls.

    For a super-operator generator L the abstract H is ``i L``
    (mu.py:129-132), hence the extra factor i.

    Raises:
        NotImplementedError: the pulse drives a collapse operator.
    """
    objective = objectives[i_objective]
    positions = pulses_mapping[i_objective][0][i_pulse]
    for i_c_op in range(len(objective.c_ops)):
        if len(pulses_mapping[i_objective][i_c_op + 1][i_pulse]) != 0:
            raise NotImplementedError(
                "Time-dependent collapse operators not implemented"
            )
    if len(positions) == 0:
        return lambda state: 0 * state
    ops = [objective.H[i][0] for i in positions]
    total = ops[0]
    for op in ops[1:]:
        total = total + op
    if kind_of(ops[0]) == 'super':
        total = 1j * total
    return total
