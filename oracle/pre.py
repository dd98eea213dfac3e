"""Oracle for the pre-processors (test infrastructure).  Reference: ``pre.py:90-149``."""

import numpy as np


def preemphasize(signal, coeff=0.97):
    """``y[0] = x[0]; y[i] = x[i] - coeff * x[i-1]`` in float64 (pre.py:136-149)"""
    out = np.asarray(signal, dtype=np.float64).copy()
    out[1:] -= coeff * np.asarray(signal, dtype=np.float64)[:-1]
    return out
