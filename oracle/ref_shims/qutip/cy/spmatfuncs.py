"""Shim (test infrastructure only) for
/root/reference/src/krotov/propagators.py:72: ``out += a * (CSR @ vec)``."""
import numpy as np


def spmvpy_csr(data, ind, ptr, vec, a, out):
    nrows = len(ptr) - 1
    for r in range(nrows):
        s = 0j
        for j in range(ptr[r], ptr[r + 1]):
            s += data[j] * vec[ind[j]]
        out[r] += a * s
