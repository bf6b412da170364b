"""Shim of ``qutip.superoperator.mat2vec/vec2mat`` (column-stacking), used at
/root/reference/src/krotov/propagators.py:73 (test infrastructure only)."""
import numpy as np


def mat2vec(mat):
    return np.asarray(mat).T.reshape(np.prod(np.shape(mat)), 1)


def vec2mat(vec):
    n = int(np.sqrt(len(vec)))
    return np.asarray(vec).reshape((n, n)).T
