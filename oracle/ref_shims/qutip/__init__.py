"""TEST INFRASTRUCTURE ONLY -- a dense, numpy-backed stand-in for the slice of
QuTiP 4.x (pinned by the reference at ``qutip>=4.3.1,<5.0``,
/root/reference/pyproject.toml:33; recorded outputs used 4.7.6,
/root/reference/binder/environment.yml:6) that the reference's Krotov hot path
touches.  QuTiP is not installable in this image (no wheel, no network), so
this shim lets the *unmodified* reference package be imported from
/root/reference/src in order to generate golden vectors
(oracle/make_golden.py) and lets tests feed Qobj-shaped objects to the
krotov_b200 facade.  It is never imported by the product package.

Semantics restated from QuTiP 4.7's published behaviour for dense data:

* ``Qobj.type`` from ``dims`` (ket / bra / oper / super),
* ``Qobj.expm()`` == ``scipy.linalg.expm`` of the dense matrix,
* ``Qobj.__call__``: ``A*psi`` for oper-on-ket; for super-on-oper
  ``vec^-1(A @ vec(rho))`` with column-stacking ``vec``,
* ``Qobj.overlap``: ``<a|b>`` for kets, ``tr(a^dag b)`` for operators,
* ``Qobj.norm()``: L2 for kets/bras, trace norm for operators,
* ``isherm``: ``|A - A^dag| <= 1e-12`` elementwise (``settings.atol``),
* ``liouvillian(H, c_ops)``:
  ``-i(I (x) H - H^T (x) I) + sum_C [C* (x) C - 1/2 I (x) C^dag C - 1/2 (C^dag C)^T (x) I]``.

Not modelled: sparse storage, ``auto_tidyup`` (drops |x|<1e-12 entries after
arithmetic), ``mesolve``.  The reference's own cross-version reproducibility
floor (scipy 1.12 vs 1.18 ``expm``) is ~1e-12 after two iterations, which is
far above any tidyup effect for the problems used here.
"""
import numbers

import numpy as np
import scipy.linalg

from . import parallel, solver, superoperator  # noqa: F401
from .superoperator import mat2vec, vec2mat  # noqa: F401

__version__ = "4.7.6+shim"

_ATOL = 1e-12


def _infer_dims(arr):
    r, c = arr.shape
    return [[r], [c]]


class Qobj:
    """Dense numpy-backed quantum object (see module docstring)."""

    __array_priority__ = 100

    def __init__(self, inpt=None, dims=None, shape=None, type=None,
                 isherm=None, copy=True, fast=False, superrep=None,
                 isunitary=None):
        if inpt is None:
            arr = np.zeros((1, 1), dtype=np.complex128)
        elif isinstance(inpt, Qobj):
            arr = inpt._data.copy()
            if dims is None:
                dims = inpt.dims
        else:
            arr = np.array(inpt, dtype=np.complex128)
            if arr.ndim == 0:
                arr = arr.reshape(1, 1)
            elif arr.ndim == 1:
                arr = arr.reshape(-1, 1)
        self._data = arr
        self.dims = dims if dims is not None else _infer_dims(arr)
        self._isherm = isherm
        self.superrep = superrep

    # ---- structure -------------------------------------------------------
    @property
    def shape(self):
        return self._data.shape

    @property
    def type(self):
        d = self.dims
        if isinstance(d[0][0], list):
            return 'super'
        r, c = self._data.shape
        if c == 1 and r > 1:
            return 'ket'
        if r == 1 and c > 1:
            return 'bra'
        return 'oper'

    @property
    def isherm(self):
        if self._isherm is not None:
            return self._isherm
        a = self._data
        if a.shape[0] != a.shape[1]:
            return False
        return bool(np.all(np.abs(a - a.conj().T) <= _ATOL))

    @isherm.setter
    def isherm(self, v):
        self._isherm = v

    @property
    def data(self):
        import scipy.sparse
        return scipy.sparse.csr_matrix(self._data)

    def full(self, order='C', squeeze=False):
        out = np.array(self._data, order=order)
        return out.squeeze() if squeeze else out

    def copy(self):
        return Qobj(self._data.copy(), dims=[list(self.dims[0]),
                                             list(self.dims[1])],
                    isherm=self._isherm)

    # ---- arithmetic ------------------------------------------------------
    def _wrap(self, arr, dims=None):
        return Qobj(arr, dims=dims if dims is not None else self.dims)

    def __add__(self, other):
        if isinstance(other, Qobj):
            return self._wrap(self._data + other._data)
        if isinstance(other, numbers.Number) or np.isscalar(other):
            if other == 0:
                return self.copy()
            return self._wrap(
                self._data + other * np.eye(*self._data.shape))
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-1) * other

    def __rsub__(self, other):
        return (-1) * self + other

    def __neg__(self):
        return self._wrap(-self._data)

    def __mul__(self, other):
        if isinstance(other, Qobj):
            out = self._data @ other._data
            dims = [self.dims[0], other.dims[1]]
            return Qobj(out, dims=dims)
        if isinstance(other, numbers.Number) or np.isscalar(other):
            return self._wrap(self._data * other)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, numbers.Number) or np.isscalar(other):
            return self._wrap(other * self._data)
        return NotImplemented

    def __truediv__(self, other):
        if isinstance(other, numbers.Number) or np.isscalar(other):
            return self._wrap(self._data / other)
        return NotImplemented

    def __eq__(self, other):
        if not isinstance(other, Qobj):
            return False
        if self._data.shape != other._data.shape:
            return False
        return bool(np.all(np.abs(self._data - other._data) <= _ATOL))

    def __ne__(self, other):
        return not self == other

    __hash__ = None

    def __getitem__(self, ind):
        return self._data[ind]

    # ---- linear algebra --------------------------------------------------
    def dag(self):
        return Qobj(self._data.conj().T, dims=[self.dims[1], self.dims[0]])

    def conj(self):
        return self._wrap(self._data.conj())

    def trans(self):
        return Qobj(self._data.T, dims=[self.dims[1], self.dims[0]])

    def tr(self):
        t = np.trace(self._data)
        return float(t.real) if self.isherm else complex(t)

    def norm(self, norm=None, sparse=False, tol=0, maxiter=100000):
        a = self._data
        if self.type in ('ket', 'bra'):
            if norm in (None, 'l2'):
                return float(np.linalg.norm(a))
            if norm == 'max':
                return float(np.max(np.abs(a)))
            raise ValueError(norm)
        if norm in (None, 'tr'):
            return float(np.sum(scipy.linalg.svdvals(a)))
        if norm == 'fro':
            return float(np.linalg.norm(a))
        if norm == 'one':
            return float(np.linalg.norm(a, 1))
        if norm == 'max':
            return float(np.max(np.abs(a)))
        raise ValueError(norm)

    def expm(self, method='dense'):
        return self._wrap(scipy.linalg.expm(self._data))

    def overlap(self, other):
        a, b = self, other
        ta, tb = a.type, b.type
        if ta == 'ket' and tb == 'ket':
            return complex((a._data.conj().T @ b._data)[0, 0])
        if ta == 'bra' and tb == 'ket':
            return complex((a._data @ b._data)[0, 0])
        if ta == 'ket' and tb == 'bra':
            return complex((a._data.conj().T @ b._data.conj().T)[0, 0])
        if ta == 'bra' and tb == 'bra':
            return complex((a._data @ b._data.conj().T)[0, 0])
        if ta == 'oper' and tb == 'oper':
            return complex(np.trace(a._data.conj().T @ b._data))
        raise TypeError("Can only calculate overlap for state vector Qobjs")

    def __call__(self, other):
        if not isinstance(other, Qobj):
            raise TypeError("Only defined for quantum objects.")
        if self.type == 'oper':
            if other.type == 'ket':
                return self * other
            raise TypeError("Can only act oper on ket.")
        if self.type == 'super':
            if other.type == 'ket':
                other = ket2dm(other)
            if other.type == 'oper':
                d = other._data.shape[0]
                v = other._data.reshape(-1, order='F')
                out = (self._data @ v).reshape(d, d, order='F')
                return Qobj(out, dims=other.dims)
            raise TypeError("Can only act super on oper or ket.")
        raise TypeError("Invalid type for __call__")

    def __repr__(self):
        return "Qobj(type=%s, shape=%s)" % (self.type, self.shape)


# ---- constructors / helpers ---------------------------------------------

def ket(seq, dim=2):
    if isinstance(seq, str):
        seq = [int(c) for c in seq]
    n = len(seq)
    dims = [dim] * n if isinstance(dim, int) else list(dim)
    idx = 0
    for s, d in zip(seq, dims):
        idx = idx * d + s
    v = np.zeros((int(np.prod(dims)), 1), dtype=np.complex128)
    v[idx, 0] = 1
    return Qobj(v, dims=[dims, [1] * n])


def basis(N, n=0):
    v = np.zeros((N, 1), dtype=np.complex128)
    v[n, 0] = 1
    return Qobj(v)


def ket2dm(psi):
    return Qobj(psi._data @ psi._data.conj().T,
                dims=[psi.dims[0], psi.dims[0]])


def identity(N):
    return Qobj(np.eye(N))


qeye = identity


def sigmax():
    return Qobj([[0, 1], [1, 0]])


def sigmay():
    return Qobj([[0, -1j], [1j, 0]])


def sigmaz():
    return Qobj([[1, 0], [0, -1]])


def sigmap():
    return Qobj([[0, 1], [0, 0]])


def sigmam():
    return Qobj([[0, 0], [1, 0]])


def tensor(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    out = args[0]._data
    d0 = list(args[0].dims[0])
    d1 = list(args[0].dims[1])
    for q in args[1:]:
        out = np.kron(out, q._data)
        d0 += list(q.dims[0])
        d1 += list(q.dims[1])
    return Qobj(out, dims=[d0, d1])


def expect(oper, state):
    if state.type == 'ket':
        v = (state._data.conj().T @ oper._data @ state._data)[0, 0]
    else:
        v = np.trace(oper._data @ state._data)
    return float(v.real) if oper.isherm else complex(v)


def liouvillian(H=None, c_ops=None, data_only=False, chi=None):
    """Dense restatement of ``qutip.liouvillian`` (column-stacking vec)."""
    if c_ops is None:
        c_ops = []
    if H is not None and H.type == 'super':
        L = H._data.copy()
        d = int(round(np.sqrt(L.shape[0])))
        hd = H.dims[0][0]
    else:
        d = (H if H is not None else c_ops[0])._data.shape[0]
        hd = (H if H is not None else c_ops[0]).dims[0]
        L = np.zeros((d * d, d * d), dtype=np.complex128)
        if H is not None:
            I = np.eye(d)
            L += -1j * (np.kron(I, H._data) - np.kron(H._data.T, I))
    I = np.eye(d)
    for c in c_ops:
        if c.type == 'super':
            L += c._data
            continue
        C = c._data
        cdc = C.conj().T @ C
        L += np.kron(C.conj(), C)
        L += -0.5 * np.kron(I, cdc)
        L += -0.5 * np.kron(cdc.T, I)
    return Qobj(L, dims=[[hd, hd], [hd, hd]])


def operator_to_vector(op):
    v = op._data.reshape(-1, 1, order='F')
    return Qobj(v, dims=[[op.dims[0], op.dims[1]], [1]])


def vector_to_operator(vec):
    n = int(round(np.sqrt(vec._data.shape[0])))
    return Qobj(vec._data.reshape(n, n, order='F'))


def mesolve(*args, **kwargs):
    raise NotImplementedError("qutip.mesolve is not available in the shim")


class _Operators:
    sigmax = staticmethod(sigmax)
    sigmay = staticmethod(sigmay)
    sigmaz = staticmethod(sigmaz)
    sigmap = staticmethod(sigmap)
    sigmam = staticmethod(sigmam)
    identity = staticmethod(identity)
    qeye = staticmethod(identity)


operators = _Operators()
