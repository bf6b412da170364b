"""Shim of ``qutip.parallel`` (test infrastructure only): ``serial_map`` with
the QuTiP 4.x signature used at /root/reference/src/krotov/optimize.py:10."""


def serial_map(task, values, task_args=tuple(), task_kwargs={}, **kwargs):
    return [task(value, *task_args, **task_kwargs) for value in values]


def parallel_map(task, values, task_args=tuple(), task_kwargs={}, **kwargs):
    return serial_map(task, values, task_args, task_kwargs)
