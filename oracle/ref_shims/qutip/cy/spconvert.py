"""Shim (test infrastructure only) for
/root/reference/src/krotov/propagators.py:71."""
import numpy as np


def dense2D_to_fastcsr_fmode(mat, nrows, ncols):
    return np.array(mat, dtype=np.complex128).reshape(nrows, ncols)
