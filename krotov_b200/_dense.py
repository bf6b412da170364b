"""Duck-typed access to quantum objects.

The reference hands ``qutip.Qobj`` instances (or, with
``Objective.type_checking = False``, plain numpy arrays -- notebook
``09_example_numpy.ipynb``) to its plugins.  The B200 engine only needs dense
complex128 data plus the object *kind*, so everything that enters the problem
compiler goes through :func:`dense` / :func:`kind_of`, and everything handed
back to user callbacks goes through :func:`like` so that hooks see the same
kind of object they supplied (ndarray in, ndarray out; Qobj in, Qobj out when
the object's class can be re-instantiated with ``cls(array, dims=...)``).
"""
import numpy as np

__all__ = ['dense', 'kind_of', 'like', 'is_quantum_object', 'adjoint_of']


def is_quantum_object(x):
    """True for ndarray or Qobj-like (has ``full()`` and ``dims``)."""
    return isinstance(x, np.ndarray) or (
        hasattr(x, 'full') and hasattr(x, 'dims')
    )


def dense(x):
    """Dense complex128 2-D array of `x` (kets become column vectors)."""
    if hasattr(x, 'full') and callable(x.full):
        a = np.asarray(x.full(), dtype=np.complex128)
    else:
        a = np.asarray(x, dtype=np.complex128)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError("quantum object must be 1-D or 2-D, not %d-D" % a.ndim)
    return a


def kind_of(x):
    """'ket', 'bra', 'oper' or 'super'.

    Qobj-likes report their own ``type``; arrays are classified by shape
    (a square array is 'oper' -- whether it acts as a superoperator is decided
    by the problem compiler from the state it is applied to).
    """
    t = getattr(x, 'type', None)
    if isinstance(t, str) and t in ('ket', 'bra', 'oper', 'super'):
        return t
    a = np.asarray(x)
    if a.ndim == 1:
        return 'ket'
    r, c = a.shape
    if c == 1 and r > 1:
        return 'ket'
    if r == 1 and c > 1:
        return 'bra'
    return 'oper'


def like(template, array):
    """Wrap `array` as the same kind of object as `template`."""
    if isinstance(template, np.ndarray) or not hasattr(template, 'full'):
        a = np.asarray(array)
        t = np.asarray(template)
        return a.reshape(t.shape) if a.size == t.size else a
    try:
        return type(template)(array, dims=template.dims)
    except Exception:  # pragma: no cover - exotic Qobj-likes
        return array


def adjoint_of(x):
    """Conjugate transpose of a Qobj-like (``dag()``) or array-like."""
    if hasattr(x, 'dag') and callable(x.dag):
        return x.dag()
    if hasattr(x, 'conj'):
        return x.conj().T
    return x.conjugate().transpose()
